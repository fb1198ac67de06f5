"""GPU: the stage-2/3 pieces (BASELINE configs[2]) against fixtures produced by the unmodified reference
(oracle/make_golden.py) and against the oracle at the bench shape."""
import pytest
import torch

from oracle import nefes_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def nb():
    import nefes_b200
    from nefes_b200 import _lib
    _lib.lib()
    return nefes_b200


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("l1", [True, False])
def test_stage23_loss_kernels_vs_reference_fixture(nb, golden, l1):
    """nefes_feat_loss_{fwd,bwd} + nefes_nerfw_loss_*: ColorFeatureFusionNerfWLoss in its three call modes; values 2e-6,
    gradients of the caller's weighted sum (run_nefes.py:238-251: 1, 0.04, 0.02) 1e-5, sign(0) = 0 kept."""
    g = golden("g6_loss.npz")
    tg = {k[7:]: v.to(DEV) for k, v in g.items() if k.startswith("target/")}
    lf = nb.ColorFeatureFusionNerfWLoss(coef=1, L1_loss=l1)
    name = "l1" if l1 else "mse"
    for tag, kw in (("color", dict(switch_on=False, color_only_switch=True)), ("stage2", dict(switch_on=False, color_only_switch=False)),
                    ("stage3", dict(switch_on=True, color_only_switch=False))):
        leaf = {k[3:]: v.to(DEV).clone().requires_grad_(True) for k, v in g.items() if k.startswith("in/")}
        out = lf(leaf, tg, **kw)
        out = out if isinstance(out, tuple) else (out,)
        for i, v in enumerate(out):
            assert rel(v, g[f"{name}/{tag}/{i}"]) < 2e-6, (name, tag, i)
        sum(w * v for w, v in zip((1.0, 0.04, 0.02), out)).backward()
        for k in ("feat_fine", "feat_coarse", "feat_fusion"):
            key = f"{name}/{tag}/grad/{k}"
            if key in g:
                assert rel(leaf[k].grad, g[key]) < 1e-5, key
                if l1 and k == "feat_fine":
                    assert float(leaf[k].grad[:4].abs().max()) == 0.0      # a == t exactly: torch's sign(0) = 0
            else:
                assert leaf[k].grad is None, key


def test_feat_loss_error_paths(nb):
    from nefes_b200.losses import _FeatLoss
    a = torch.zeros(8, 128, device=DEV)
    with pytest.raises(RuntimeError):
        _FeatLoss.apply(a, None, torch.zeros(8, 64, device=DEV), 0)
    with pytest.raises(RuntimeError):
        _FeatLoss.apply(a.cpu(), None, a.cpu(), 0)
