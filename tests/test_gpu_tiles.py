"""Tile-major raw layout (NEFES_RAW_TILES): the fused chain + tile-major compositing must give the same numbers as
the row-major public API of the same kernels (bit-exact: identical arithmetic, different addresses)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nets():
    import nefes_b200 as nb
    c = nb.NeRFH_NFF("coarse", W=128).cuda()
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True).cuda()
    c.precision = f.precision = "bf16"
    return c, f


@pytest.mark.parametrize("n_rays,S,net", [(96, 128, "fine"), (67, 64, "coarse"), (66, 64, "coarse"), (5, 32, "fine")])
def test_tiled_query_and_composite_match_rows(n_rays, S, net):
    from nefes_b200 import _lib as L, ops
    from nefes_b200.nerfh_nff import raw2outputs_NeRFH_NFF
    c, f = _nets()
    m = f if net == "fine" else c
    mode = L.MODE_FULL if net == "fine" else L.MODE_STATIC
    g = torch.Generator(device="cuda").manual_seed(1)
    pts = (torch.rand(n_rays, S, 3, device="cuda", generator=g) * 4 - 2).requires_grad_(True)
    dirs = torch.nn.functional.normalize(torch.randn(n_rays, 3, device="cuda", generator=g), dim=-1)
    z = torch.sort(torch.rand(n_rays, S, device="cuda", generator=g) * 4, dim=-1).values

    def run(tiled):
        m.zero_grad()
        if pts.grad is not None:
            pts.grad = None
        if tiled:
            with ops.tiled_raw():
                raw = m.query(pts, dirs, mode)
            assert isinstance(raw, ops.TiledRaw)
        else:
            raw = m.query(pts, dirs, mode)
        out = raw2outputs_NeRFH_NFF(raw, z, output_transient=(net == "fine"), beta_min=0.1, typ=net)
        rgb, feat, disp, acc, weights, depth, tsig, beta = out
        loss = rgb.sum() + (feat * feat).sum() + depth.sum() + acc.sum() + beta.sum() + (0 if tsig is None else tsig.sum())
        loss.backward()
        rows = raw.rows() if tiled else raw
        return [rows.detach(), rgb.detach(), feat.detach(), depth.detach(), weights.detach(), m.flat.grad.clone(), pts.grad.clone()]

    a, b = run(False), run(True)
    names = ["raw", "rgb", "feat", "depth", "weights", "d_params", "d_pts"]
    for name, x, y in zip(names, a, b):
        if name in ("raw", "weights"):
            assert torch.equal(x, y), name
        else:
            # forward reductions run in a different order (thread-per-channel vs warp-per-channel): 1e-6 relative; the
            # gradients are then re-rounded to bf16 images, where a 1-ulp input change is a 4e-3 output change
            tol = 1e-2 if name.startswith("d_") else 2e-5
            scale = float(x.abs().max()) + 1e-12
            assert float((x - y).abs().max()) <= tol * scale, (name, float((x - y).abs().max()), scale)


def test_nerfw_loss_matches_torch_expression():
    """NerfWLoss kernels (losses.py:96-132) against the reference's torch expression, value and all four gradients."""
    import nefes_b200 as nb
    g = torch.Generator(device="cuda").manual_seed(3)
    N, S = 777, 128
    rgb0 = torch.rand(N, 3, device="cuda", generator=g).requires_grad_(True)
    rgb = torch.rand(N, 3, device="cuda", generator=g).requires_grad_(True)
    beta = (torch.rand(N, device="cuda", generator=g) + 0.1).requires_grad_(True)
    tsig = torch.rand(N, S, device="cuda", generator=g).requires_grad_(True)
    tgt = torch.rand(N, 3, device="cuda", generator=g)

    def ref():
        c_l = 0.5 * ((rgb0 - tgt) ** 2).mean()
        f_l = ((rgb - tgt) ** 2 / (2 * beta.unsqueeze(1) ** 2)).mean()
        return 2.0 * (c_l + f_l + 3 + torch.log(beta).mean() + 0.02 * tsig.mean())

    lr = ref()
    gr = torch.autograd.grad(lr * 1.5, [rgb0, rgb, beta, tsig])
    le = nb.NerfWLoss(coef=2.0, lambda_u=0.02)({"rgb_coarse": rgb0, "rgb_fine": rgb, "beta": beta, "transient_sigmas": tsig}, tgt)
    ge = torch.autograd.grad(le * 1.5, [rgb0, rgb, beta, tsig])
    assert abs(float(le) - float(lr)) <= 2e-6 * abs(float(lr))
    for a, b in zip(ge, gr):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12
