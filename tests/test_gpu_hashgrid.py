"""GPU: encoder front-end B (HashGrid + SH) through the C ABI against oracle/hashgrid_oracle.py --
a restatement of tiny-cuda-nn's published algorithm (PARITY UNPINNED: no tcnn here, dead code in the reference)."""
import pytest
import torch

from oracle import hashgrid_oracle as HO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def levels_of(enc):
    """The restatement's level table with the scale values taken bit-for-bit from the library (tiny-cuda-nn and
    the engine both evaluate exp2f on the host in C; numpy's float32 exp2 can differ in the last ulp, which at a
    grid position of ~2000 cells moves the trilinear weights by 1e-4)."""
    levels, total = HO.hash_layout()
    for l, lv in enumerate(levels):
        lv["scale"] = float(enc.layout.level[l].scale)
    return levels, total


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_hash_layout_matches_restatement():
    from nefes_b200.hashgrid import HashGridEncoding
    enc = HashGridEncoding()
    levels, total = HO.hash_layout()
    assert enc.layout.n_entries == total == 6098120            # SURVEY 8c arithmetic
    for l, lv in enumerate(levels):
        got = enc.layout.level[l]
        assert (got.res, got.size, got.offset, bool(got.dense)) == (lv["res"], lv["size"], lv["offset"], lv["dense"])
        assert abs(got.scale - lv["scale"]) <= 1e-6 * lv["scale"]


def test_hash_encode_forward_backward():
    from nefes_b200.hashgrid import HashGridEncoding
    gen = torch.Generator().manual_seed(0)
    enc = HashGridEncoding().to(DEV)
    levels, total = levels_of(enc)
    table = (torch.rand(total, 2, generator=gen) * 2 - 1)       # O(1) entries so errors are visible
    with torch.no_grad():
        enc.params.copy_(table.reshape(-1).to(DEV))
    x = torch.rand(3000, 3, generator=gen)
    x[:8] = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0., 1., 0.5], [1., 0., 0.], [.5, .5, .5], [0.999999, 0.5, 0.], [0, 0, 1], [1, 1, 0]])
    xg = x.to(DEV).requires_grad_(True)
    out = enc(xg)
    xr, tr = x.clone().requires_grad_(True), table.clone().requires_grad_(True)
    ref = HO.hash_encode(xr, tr, levels)
    assert out.shape == ref.shape == (3000, 32)
    assert rel(out, ref) < 5e-6
    k = torch.randn(ref.shape, generator=gen)
    (out * k.to(DEV)).sum().backward()
    (ref * k).sum().backward()
    assert rel(enc.params.grad.reshape(-1, 2), tr.grad) < 1e-5
    assert rel(xg.grad, xr.grad) < 1e-4


def test_sh_forward_backward():
    from nefes_b200.hashgrid import SHEncoding
    gen = torch.Generator().manual_seed(1)
    d = torch.rand(2000, 3, generator=gen)
    dg = d.to(DEV).requires_grad_(True)
    out = SHEncoding()(dg)
    dr = d.clone().requires_grad_(True)
    ref = HO.sh_encode(dr)
    assert rel(out, ref) < 1e-6
    k = torch.randn(ref.shape, generator=gen)
    (out * k.to(DEV)).sum().backward()
    (ref * k).sum().backward()
    assert rel(dg.grad, dr.grad) < 1e-5


def test_tcnn_shaped_field():
    """NeRFH_TCNN.forward (coarse form) against the restated field with the same parameters."""
    from nefes_b200.hashgrid import NeRFH_TCNN
    gen = torch.Generator().manual_seed(2)
    m = NeRFH_TCNN("coarse", bound=4).to(DEV)
    with torch.no_grad():
        m.encoder.params.copy_(((torch.rand(m.encoder.params.shape, generator=gen) * 2 - 1) * 0.5).to(DEV))
    levels, _ = levels_of(m.encoder)
    P = {"table": m.encoder.params.detach().cpu().reshape(-1, 2), "sigma.0": m.sigma_net[0].detach().cpu(),
         "sigma.1": m.sigma_net[1].detach().cpu(), "color.0": m.color_net[0].detach().cpu(),
         "color.1": m.color_net[1].detach().cpu(), "color.2": m.color_net[2].detach().cpu()}
    x = torch.rand(1500, 3, generator=gen) * 8 - 4
    d = torch.nn.functional.normalize(torch.randn(1500, 3, generator=gen), dim=-1)
    out = m(x.to(DEV), d.to(DEV))
    ref = HO.tcnn_field_forward(P, x, d, bound=4.0, levels=levels)
    assert out.shape == (1500, 4) and rel(out, ref) < 1e-4
    sig = m(x.to(DEV), d.to(DEV), sigma_only=True)
    assert rel(sig, HO.tcnn_field_forward(P, x, d, bound=4.0, levels=levels, sigma_only=True)) < 1e-4
    out.sum().backward()
    assert m.encoder.params.grad is not None and torch.isfinite(m.encoder.params.grad).all()
    fine = NeRFH_TCNN("fine", encode_appearance=True, encode_transient=True, bound=4).to(DEV)
    ts = torch.zeros(1500, 10, device=DEV)
    assert fine(x.to(DEV), d.to(DEV), ts=ts, output_transient=True).shape == (1500, 9)
