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
    fine = NeRFH_TCNN("fine", encode_appearance=True, encode_transient=True, in_channels_a=50, in_channels_t=20, bound=4).to(DEV)
    ts = torch.zeros(1500, 10, device=DEV)
    assert fine(x.to(DEV), d.to(DEV), ts=ts, output_transient=True).shape == (1500, 9)


def test_front_end_b_through_render_rays():
    """Front-end B ON THE RENDER PATH: render() with NeRFH_TCNN fields and the tcnn query function (run_NeRFH_TCNN,
    nerfh_tcnn.py:368-440) -- stratified depths, hash / SH encodings, heads, compositing of the 4- and 9-channel raw,
    sample_pdf, merge -- in train mode with the NeRF-W transient head, against the same pipeline restated on the CPU
    (oracle/hashgrid_oracle.py field + nefes_oracle compositing / sampling; PARITY UNPINNED for the tiny-cuda-nn part).
    Also: weight and table gradients arrive, and the test-time (sigma-only coarse pass) route runs."""
    import nefes_b200 as nb
    from nefes_b200.hashgrid import NeRFH_TCNN, TcnnQuery
    from oracle import nefes_oracle as O
    gen = torch.Generator().manual_seed(4)
    coarse = NeRFH_TCNN("coarse", bound=4).to(DEV)
    fine = NeRFH_TCNN("fine", encode_appearance=True, encode_transient=True, in_channels_a=50, in_channels_t=20, bound=4).to(DEV)
    with torch.no_grad():
        for m in (coarse, fine):                         # O(1) table entries, otherwise the field is numerically flat
            m.encoder.params.copy_(((torch.rand(m.encoder.params.shape, generator=gen) * 2 - 1) * 0.5).to(DEV))
    levels, _ = levels_of(coarse.encoder)

    def params(m, fine_net):
        P = {"table": m.encoder.params.detach().cpu().reshape(-1, 2), "sigma.0": m.sigma_net[0].detach().cpu(),
             "sigma.1": m.sigma_net[1].detach().cpu(), "color.0": m.color_net[0].detach().cpu(),
             "color.1": m.color_net[1].detach().cpu(), "color.2": m.color_net[2].detach().cpu()}
        if fine_net:
            P.update({"emb_a": m.embedding_a.weight.detach().cpu(), "emb_t": m.embedding_t.weight.detach().cpu()})
            P.update({f"trans.{i}": w.detach().cpu() for i, w in enumerate(m.transient_color_net)})
        return P
    Pc, Pf = params(coarse, False), params(fine, True)
    n, H, W, focal = 96, 60, 80, 65.7
    pose = torch.eye(4)[:3]
    ro, rd = O.camera_rays(H, W, focal, pose)
    pix = torch.randperm(H * W, generator=gen)[:n]
    ro, rd = ro.reshape(-1, 3)[pix].contiguous(), rd.reshape(-1, 3)[pix].contiguous()
    hist = torch.randint(0, 30, (1, 10), generator=gen).float()
    t_rand, u = torch.rand(n, 64, generator=gen), torch.rand(n, 64, generator=gen)

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test = False, False, True, True
    kw = dict(network_query_fn=TcnnQuery(1 << 16), N_importance=64, N_samples=64, network_fn=coarse, network_fine=fine, use_viewdirs=True,
              white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0.5, far=4., raw_noise_std=0.)
    rgb, disp, acc, ex = nb.render(H, W, focal, rays=(ro.to(DEV), rd.to(DEV)), img_idx=hist.to(DEV), perturb=1., test_time=False,
                                   t_rand=t_rand.to(DEV), u=u.to(DEV), **kw)
    assert "feat_map" not in ex and set(ex) >= {"rgb0", "disp0", "acc0", "z_std", "transient_sigmas", "beta"}
    # ---- the same pipeline on the CPU ----
    view = rd / rd.norm(dim=-1, keepdim=True)
    z_c = O.coarse_depths(torch.full((n, 1), 0.5), torch.full((n, 1), 4.0), 64, t_rand)
    pts = ro[:, None] + rd[:, None] * z_c[..., None]
    raw_c = HO.tcnn_field_forward(Pc, pts.reshape(-1, 3), view[:, None].expand(pts.shape).reshape(-1, 3), bound=4.0, levels=levels).reshape(n, 64, 4)
    c0 = O.composite(raw_c, z_c, typ="coarse")
    z_s, _, _ = O.importance_depths(.5 * (z_c[..., 1:] + z_c[..., :-1]), c0.weights[..., 1:-1], 64, u)
    z_f, _ = torch.sort(torch.cat([z_c, z_s], -1), -1)
    pts_f = ro[:, None] + rd[:, None] * z_f[..., None]
    ts_pt = hist.expand(n, -1)[:, None, :].expand(-1, 128, -1).reshape(-1, 10)
    raw_f = HO.tcnn_field_forward_fine(Pf, pts_f.reshape(-1, 3), view[:, None].expand(pts_f.shape).reshape(-1, 3), ts_pt, bound=4.0,
                                       levels=levels).reshape(n, 128, 9)
    c1 = O.composite(raw_f, z_f, output_transient=True, beta_min=fine.beta_min, typ="fine", transient_at_test=True)
    for name, got, want in (("rgb", rgb, c1.rgb), ("acc", acc, c1.acc), ("disp", disp, c1.disp), ("rgb0", ex["rgb0"], c0.rgb),
                            ("beta", ex["beta"], c1.beta)):
        assert rel(got, want) < 2e-3, (name, rel(got, want))
    # per-SAMPLE values of a random O(1) hash table (a very rough field): a 1e-6 shift of a fine depth moves them visibly;
    # the mean is what the loss uses
    assert rel(ex["transient_sigmas"], c1.transient_sigmas) < 1e-1
    assert abs(float(ex["transient_sigmas"].mean()) - float(c1.transient_sigmas.mean())) < 2e-3 * float(c1.transient_sigmas.mean())
    (rgb.sum() + ex["rgb0"].sum() + ex["beta"].sum()).backward()
    for m in (coarse, fine):
        assert m.encoder.params.grad is not None and float(m.encoder.params.grad.abs().max()) > 0
        assert all(w.grad is not None and torch.isfinite(w.grad).all() for w in m.color_net)
    with torch.no_grad():
        rgb_t, _, _, ex_t = nb.render(H, W, focal, rays=(ro.to(DEV), rd.to(DEV)), img_idx=hist.to(DEV), perturb=0., test_time=True, **kw)
    assert rgb_t.shape == (n, 3) and torch.isfinite(rgb_t).all() and "rgb0" not in ex_t
