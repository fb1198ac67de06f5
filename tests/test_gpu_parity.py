"""GPU parity: every CUDA stage, through the C ABI (nefes_b200 -> ctypes -> libnefes_b200.so),
against (a) the committed golden vectors produced by the unmodified reference and (b) the CPU
oracle on larger seeded inputs.  Bars: bit-exact for rays / coarse depths / sample indices given
the same cdf; fp32 tolerances stated per test (north star: 1e-3 relative)."""
import math

import pytest
import torch

from oracle import nefes_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.


def pdf_tolerance(bins, cdf, inds, ulps=8):
    """Conditioning bound of the inverse-CDF interpolation: z = b_lo + (u - c_lo)/(c_hi - c_lo) * (b_hi - b_lo),
    so an `ulps`-ulp difference in the cdf knots (the normaliser is an order-dependent fp32 sum) moves z by at
    most ulps * 2^-24 * width / denom.  Flat pdf bins (tiny denom) are ill-conditioned by construction, in
    the reference as much as here; the move is always bounded by the bin width."""
    inds = inds.long()
    lo, hi = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
    denom = torch.gather(cdf, -1, hi) - torch.gather(cdf, -1, lo)
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    width = (torch.gather(bins, -1, hi) - torch.gather(bins, -1, lo)).abs()
    return 1e-6 + ulps * 2.0 ** -24 * width / denom


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def nb():
    import nefes_b200
    from nefes_b200 import _lib
    _lib.lib()
    return nefes_b200


@pytest.fixture(scope="module")
def models(nb, weights):
    wc, wf = weights
    c = nb.NeRFH_NFF("coarse", W=128)
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True)
    c.load_state_dict(wc, strict=False)
    f.load_state_dict(wf)
    return c.to(DEV), f.to(DEV)


class Args:
    nerfh_nff = True
    use_fine_only = False
    NeRFW = True
    transient_at_test = True
    netchunk = 1 << 21


def render_kwargs(nb, models, test_time, fused=False):
    """fused=False: a hand-written query closure, as the reference builds it -> the staged path (one engine call per
    stage); fused=True: nb.StandardQuery -> render_rays is ONE engine call (nefes_render_rays_fwd/_bwd)."""
    c, f = models
    q = lambda inputs, viewdirs, ts, fn, typ, output_transient, test_time, store_rgb: \
        nb.run_network_NeRFH_NFF(inputs, viewdirs, ts, fn, typ=typ, output_transient=output_transient,
                                 netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
    if fused:
        q = nb.StandardQuery(Args.netchunk)
    return dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=c, network_fine=f,
                use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR,
                perturb=0. if test_time else 1., raw_noise_std=0., test_time=test_time)


# ---------------------------------------------------------------------------------------------
def test_get_rays_bit_exact_and_pose_gradient(nb, golden):
    g = golden("g1_rays.npz")
    c2w = g["c2w"].to(DEV).requires_grad_(True)
    o, d = nb.get_rays(H, W, FOCAL, c2w)
    assert torch.equal(o.cpu(), g["rays_o"]) and torch.equal(d.detach().cpu(), g["rays_d"])
    ob, db = nb.get_rays_batch(H, W, FOCAL, g["c2w_b"].to(DEV))
    assert torch.equal(ob.cpu(), g["rays_o_b"]) and torch.equal(db.cpu(), g["rays_d_b"])
    gen = torch.Generator().manual_seed(1)
    ko, kd = torch.randn(H, W, 3, generator=gen), torch.randn(H, W, 3, generator=gen)
    ((o * ko.to(DEV)).sum() + (d * kd.to(DEV)).sum()).backward()
    c_ref = g["c2w"].clone().requires_grad_(True)
    o2, d2 = O.camera_rays(H, W, FOCAL, c_ref)
    ((o2 * ko).sum() + (d2 * kd).sum()).backward()
    assert rel_err(c2w.grad, c_ref.grad) < 1e-5


def test_sample_coarse_bit_exact(nb):
    from nefes_b200 import ops
    gen = torch.Generator().manual_seed(2)
    n = 777
    near = torch.rand(n, 1, generator=gen)
    far = near + 1 + 5 * torch.rand(n, 1, generator=gen)
    t_rand = torch.rand(n, 64, generator=gen)
    rb = torch.cat([torch.zeros(n, 6), near, far, torch.zeros(n, 13)], 1).to(DEV)
    for tr in (None, t_rand):
        z = ops.sample_coarse(rb[:, 6], rb[:, 7], 21, n, 64, None if tr is None else tr.to(DEV))
        assert torch.equal(z.cpu(), O.coarse_depths(near, far, 64, tr))


def test_sample_pdf_indices_exact_given_cdf(nb, golden):
    """Stage-level known-answer test: same (bins, cdf, u) -> identical inds and samples."""
    from nefes_b200 import ops
    g = golden("g2_sample_pdf.npz")
    bins, cdf = g["bins"].to(DEV), g["cdf"].to(DEV)
    for tag, u in (("rand", g["u_rand"]), ("pytest", g["u_pytest"]), ("det", None)):
        s, inds = ops.sample_pdf(bins, None, 64, u=None if u is None else u.to(DEV), cdf=cdf, return_inds=True)
        assert torch.equal(inds.cpu(), g["inds_" + tag]), tag
        assert torch.equal(s.cpu(), g["samples_" + tag]), tag


def test_sample_pdf_from_weights(nb, golden):
    """Full sample_pdf: the pdf normaliser is a SIMD-order-dependent torch.sum on the CPU, so the cdf
    may differ in the last ulp; indices may flip only where u sits within 2 ulp of a cdf knot."""
    from nefes_b200 import ops
    g = golden("g2_sample_pdf.npz")
    bins, wts, cdf_ref = g["bins"].to(DEV), g["weights"].to(DEV), g["cdf"]
    for tag, u in (("rand", g["u_rand"]), ("det", None)):
        s, inds = ops.sample_pdf(bins, wts, 64, u=None if u is None else u.to(DEV), return_inds=True)
        bad = inds.cpu() != g["inds_" + tag]
        if bad.any():
            uu = (g["u_det"].expand(96, 64) if u is None else u)[bad]
            rows = bad.nonzero()[:, 0]
            near_knot = (cdf_ref[rows] - uu[:, None]).abs().min(-1)[0]
            assert float(near_knot.max()) <= 3e-7
        assert float(bad.float().mean()) < 0.02
        ok = ~bad
        tol = pdf_tolerance(g["bins"], cdf_ref, g["inds_" + tag])
        assert float((((s.cpu() - g["samples_" + tag]).abs() > tol) & ok).float().mean()) < 2e-3
        assert float((s.cpu() - g["samples_" + tag]).abs().max()) < 2 * 4.0 / 63   # never leaves its (jittered) bin
    out = nb.sample_pdf(bins, wts, 64, det=False, pytest=True)          # the reference's own determinism hook
    tol = pdf_tolerance(g["bins"], cdf_ref, g["inds_pytest"])
    assert float(((out.cpu() - g["samples_pytest"]).abs() > tol).float().mean()) < 0.005


def test_sample_fine_sorted_union(nb):
    from nefes_b200 import ops
    gen = torch.Generator().manual_seed(5)
    n = 513
    zc = O.coarse_depths(torch.zeros(n, 1), 4 * torch.ones(n, 1), 64, torch.rand(n, 64, generator=gen))
    w = torch.rand(n, 64, generator=gen) ** 3
    u = torch.rand(n, 64, generator=gen)
    for uu in (u, None):
        zf, zs, inds = ops.sample_fine(zc.to(DEV), w.to(DEV), 64, None if uu is None else uu.to(DEV))
        mids = .5 * (zc[:, 1:] + zc[:, :-1])
        s_ref, i_ref, cdf = O.importance_depths(mids, w[:, 1:-1], 64, uu)
        zf_ref = torch.sort(torch.cat([zc, s_ref], -1), -1)[0]
        same = inds.cpu() == i_ref.int()
        assert float((~same).float().mean()) < 0.01
        tol = pdf_tolerance(mids, cdf, i_ref)
        # beyond the conditioning bound only the reference's own `denom < 1e-5 -> 1` switch (rendering.py:61) can
        # differ, when a bin's cdf mass sits within an ulp of 1e-5; the sample then still lies inside the same bin
        viol = ((zs.cpu() - s_ref).abs() > tol) & same
        assert float(viol.float().mean()) < 2e-3
        assert float((zs.cpu() - s_ref).abs().max()) < 2 * 4.0 / 63    # stays inside its (jittered) bin
        assert float(((zf.cpu() - zf_ref).abs() > 1e-4).float().mean()) < 5e-3
        assert bool((zf[:, 1:] >= zf[:, :-1]).all())
        # exact multiset property: z_fine is a permutation of cat(z_coarse, z_samples)
        assert torch.equal(torch.sort(torch.cat([zc.to(DEV), zs], -1), -1)[0], zf)


def test_positional_encoding(nb):
    from nefes_b200 import ops
    gen = torch.Generator().manual_seed(6)
    x = (torch.rand(1000, 3, generator=gen) * 8 - 4)
    for L_ in (10, 4):
        xg = x.to(DEV).requires_grad_(True)
        e = ops.encode_pe(xg, L_)
        xr = x.clone().double().requires_grad_(True)
        er = O.freq_encode(xr, L_)
        # |arg| reaches 2^9*4 rad: fp32 argument spacing is 2.4e-4 there; the oracle evaluates sin/cos of the SAME fp32
        # argument (x * 2^l is exact) in fp64 -- the host's vectorised fp32 sinf is itself only accurate to 1e-4 at
        # such arguments on some CPUs, so it cannot be the yardstick for the device's full-range sincosf.
        assert float((e.detach().cpu().double() - er.detach()).abs().max()) < 2e-6
        k = torch.randn(er.shape, generator=gen)
        (e * k.to(DEV)).sum().backward()
        (er * k.double()).sum().backward()
        assert rel_err(xg.grad, xr.grad.float()) < 1e-5


CASES = {
    "coarse_train": dict(typ="coarse", test_time=False),
    "fine_train": dict(typ="fine", test_time=False, output_transient=True, transient_at_test=True),
    "fine_test_tat": dict(typ="fine", test_time=True, output_transient=True, transient_at_test=True),
    "fine_test_static": dict(typ="fine", test_time=True, output_transient=True, transient_at_test=False),
    "fine_notransient": dict(typ="fine", test_time=False),
}
NAMES = ("rgb", "feat", "disp", "acc", "weights", "depth", "transient_sigmas", "beta")


def test_composite_golden_all_modes(nb, golden):
    g = golden("g3_composite.npz")
    for case, kw in CASES.items():
        raw = g[case + "/raw"].to(DEV).requires_grad_(True)
        z = (g["z64"] if raw.shape[1] == 64 else g["z128"]).to(DEV)
        out = nb.raw2outputs_NeRFH_NFF(raw, z, raw_noise_std=0, **kw)
        gg = torch.Generator().manual_seed(17)
        loss = 0
        for name, t in zip(NAMES, out):
            key = f"{case}/{name}"
            if key in g:
                assert rel_err(t, g[key]) < 2e-6, key
            ref_present = key in g
            if t is not None and t.requires_grad:
                assert ref_present or name == "beta"
                loss = loss + (t * torch.randn(t.shape, generator=gg).to(DEV)).sum()
        loss.backward()
        assert rel_err(raw.grad, g[case + "/d_raw"]) < 2e-5, case
    acc, w = nb.raw2outputs_NeRFH_NFF(g["coarse_test/raw"].to(DEV), g["z64"].to(DEV), typ="coarse", test_time=True)[3:5]
    assert rel_err(w, g["coarse_test/weights"]) < 2e-6 and rel_err(acc, g["coarse_test/acc"]) < 2e-6


def test_composite_vs_oracle_large(nb):
    gen = torch.Generator().manual_seed(8)
    n = 300
    z = torch.sort(torch.rand(n, 128, generator=gen) * 4, -1)[0]
    raw = torch.randn(n, 128, 137, generator=gen)
    raw[..., 131] = torch.nn.functional.softplus(raw[..., 131] * 4)
    raw[..., 132:135] = torch.sigmoid(raw[..., 132:135])
    raw[..., 135:137] = torch.nn.functional.softplus(raw[..., 135:137])
    raw[:10, :, 131] = 0                                       # empty rays: acc = 0
    raw[10:20, 5, 131] = 1e4                                   # opaque wall: transmittance hits 0
    rg = raw.to(DEV).requires_grad_(True)
    out = nb.raw2outputs_NeRFH_NFF(rg, z.to(DEV), output_transient=True, typ="fine", transient_at_test=True)
    rr = raw.clone().requires_grad_(True)
    ref = O.composite(rr, z, output_transient=True, typ="fine", transient_at_test=True).astuple()
    k = [torch.randn(t.shape, generator=gen) for t in ref]
    sum((a * b.to(DEV)).sum() for a, b in zip(out, k)).backward()
    sum((a * b).sum() for a, b in zip(ref, k)).backward()
    for name, a, b in zip(NAMES, out, ref):
        if name == "disp":      # 1/max(1e-10, depth/acc): 0/0 on empty rays is NaN on both sides
            m = torch.isfinite(b)
            assert rel_err(a.cpu()[m], b[m]) < 1e-4
            continue
        assert rel_err(a, b) < 5e-6, name
    gm = torch.isfinite(rr.grad).all(-1).all(-1)
    assert rel_err(rg.grad.cpu()[gm], rr.grad[gm]) < 5e-5


def test_mlp_forward_golden(nb, golden, models):
    g = golden("g4_mlp.npz")
    c, f = models
    emb = g["emb"].to(DEV)
    with torch.no_grad():
        assert rel_err(c(emb[:, :63], sigma_only=True), g["sigma"]) < 2e-5
        assert rel_err(c(emb, output_transient=False), g["static"]) < 2e-5
        assert rel_err(f(emb, output_transient=True), g["full"]) < 2e-5


def test_mlp_backward_vs_oracle(nb, weights, models):
    """Gradients to every parameter, to the sample positions and to the view directions."""
    wc, wf = weights
    c, f = models
    gen = torch.Generator().manual_seed(10)
    n, s = 40, 16
    pts = torch.rand(n, s, 3, generator=gen) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    for model, P, mode, typ, tr in ((f, wf, 2, "fine", True), (c, wc, 1, "coarse", False), (c, wc, 0, "coarse", False)):
        model.zero_grad()
        pg, dg = pts.to(DEV).requires_grad_(True), dirs.to(DEV).requires_grad_(True)
        raw = model.query(pg, dg, mode)
        Pg = O.clone_params(P, requires_grad=True)
        pr, dr = pts.clone().requires_grad_(True), dirs.clone().requires_grad_(True)
        ref = O.query_field(Pg, pr, dr, typ, tr, test_time=(mode == 0))
        assert rel_err(raw, ref) < 2e-5
        k = torch.randn(ref.shape, generator=gen)
        (raw * k.to(DEV)).sum().backward()
        (ref * k).sum().backward()
        assert rel_err(pg.grad, pr.grad) < 2e-4
        if mode != 0:
            assert rel_err(dg.grad, dr.grad) < 2e-4
        views = model.layer_views(model.flat.grad)
        for key, ref_g in Pg.items():
            if mode == 0 and ref_g.grad is None:
                continue
            assert rel_err(views[key], ref_g.grad) < 2e-4, key


@pytest.mark.parametrize("fused", [False, True])
def test_render_train_golden(nb, golden, models, fused):
    """End-to-end render() in train mode on the reference's RNG draws: outputs, sample indices, loss and
    weight gradients against the unmodified reference (fixture g5)."""
    g = golden("g5_render.npz")
    c, f = models
    c.zero_grad(), f.zero_grad()
    rays = (g["rays_o"].to(DEV), g["rays_d"].to(DEV))
    rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, rays=rays, img_idx=torch.zeros(1, 10),
                                   t_rand=g["train/t_rand"].to(DEV), u=g["train/u"].to(DEV), return_aux=True,
                                   retraw=True, **render_kwargs(nb, models, False, fused))
    out = dict(rgb_map=rgb, disp_map=disp, acc_map=acc, **ex)
    assert torch.equal(out["aux_z_coarse"].cpu(), g["train/z_coarse"])
    mism = (out["aux_inds"].cpu() != g["train/inds"]).float().mean()
    assert float(mism) < 0.005                                  # flips only at cdf knots
    assert float((out["aux_z_fine"].cpu() - g["train/z_fine"]).abs().max()) < 1e-4
    for k in ("rgb_map", "acc_map", "feat_map", "rgb0", "acc0", "feat0", "beta", "transient_sigmas", "disp_map", "disp0"):
        assert rel_err(out[k], g["train/" + k]) < 1e-3, k       # north-star bar; observed ~1e-5
    assert rel_err(out["z_std"], g["train/z_std"]) < 1e-3
    outn = {k: v for k, v in out.items()}
    loss = O.nerfw_loss(outn, g["train/target"].to(DEV)) + 0.04 * (out["feat_map"].abs().mean() + out["feat0"].abs().mean())
    assert abs(float(loss) - float(g["train/loss"])) < 1e-4 * abs(float(g["train/loss"]))
    loss.backward()
    vf, vc = f.layer_views(f.flat.grad), c.layer_views(c.flat.grad)
    for key, ref in g.items():
        if key.startswith("train/grad_fine/"):
            assert rel_err(vf[key[len("train/grad_fine/"):]], ref) < 2e-3, key
        if key.startswith("train/grad_coarse/"):
            assert rel_err(vc[key[len("train/grad_coarse/"):]], ref) < 2e-3, key


@pytest.mark.parametrize("fused", [False, True])
def test_render_refinement_pose_gradient_golden(nb, golden, models, fused):
    """test_time=True full-image render from c2w, cosine feature loss, gradient to the 3x4 pose."""
    g = golden("g5_render.npz")
    c, f = models
    for p in list(c.parameters()) + list(f.parameters()):
        p.requires_grad_(False)
    try:
        c2w = g["test/c2w"].to(DEV).requires_grad_(True)
        rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, c2w=c2w, img_idx=torch.zeros(1, 10), return_aux=True,
                                       **render_kwargs(nb, models, True, fused))
        sub = g["test/sub"].to(DEV)
        assert set(k for k in ex if not k.startswith("aux_")) == {"feat_map"}
        assert rel_err(ex["feat_map"][sub], g["test/feat_map"]) < 1e-3
        assert rel_err(rgb[sub], g["test/rgb_map"]) < 1e-3
        assert float((ex["aux_inds"][sub].cpu() != g["test/inds"]).float().mean()) < 0.01
        loss = O.cosine_feature_loss(ex["feat_map"][sub].t(), g["test/feat_target"].to(DEV)) + rgb[sub].mean()
        assert abs(float(loss) - float(g["test/loss"])) < 1e-4
        loss.backward()
        assert rel_err(c2w.grad, g["test/d_c2w"]) < 5e-3
    finally:
        for p in list(c.parameters()) + list(f.parameters()):
            p.requires_grad_(True)


@pytest.mark.parametrize("fused", [False, True])
def test_render_vs_oracle_1024_rays_and_properties(nb, weights, models, fused):
    """Seeded batch larger than the fixtures, checked against the oracle run on the host, plus
    size-independent properties."""
    wc, wf = weights
    gen = torch.Generator().manual_seed(12)
    n = 1024
    pose = torch.eye(4)[:3]
    o, d = O.camera_rays(H, W, FOCAL, pose)
    pix = torch.randperm(H * W, generator=gen)[:n]
    rays = (o.reshape(-1, 3)[pix] + torch.rand(n, 3, generator=gen) * 0.1, d.reshape(-1, 3)[pix])
    t_rand, u = torch.rand(n, 64, generator=gen), torch.rand(n, 64, generator=gen)
    with torch.no_grad():
        ref = O.render(H, W, FOCAL, wc, wf, rays=rays, near=NEAR, far=FAR, test_time=False, t_rand=t_rand, u=u)
        rgb, disp, acc, ex = nb.render(H, W, FOCAL, rays=(rays[0].to(DEV), rays[1].to(DEV)),
                                       img_idx=torch.zeros(1, 10), t_rand=t_rand.to(DEV), u=u.to(DEV),
                                       return_aux=True, **render_kwargs(nb, models, False, fused))
    for k, v in dict(rgb_map=rgb, acc_map=acc, feat_map=ex["feat_map"], rgb0=ex["rgb0"], feat0=ex["feat0"],
                     beta=ex["beta"]).items():
        assert rel_err(v, ref[k]) < 1e-3, k
    assert bool((ex["aux_z_fine"][:, 1:] >= ex["aux_z_fine"][:, :-1]).all())
    assert bool((acc <= 1 + 1e-5).all()) and bool((acc >= 0).all())
    assert bool((ex["aux_inds"] >= 1).all()) and bool((ex["aux_inds"] <= 63).all())
    assert bool((ex["beta"] >= 0.1).all())


def test_flat_adam_matches_torch(nb):
    from nefes_b200 import FlatAdam
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(10007, generator=gen)
    a = torch.nn.Parameter(p0.clone().to(DEV))
    b = torch.nn.Parameter(p0.clone().to(DEV))
    oa, ob = FlatAdam([a], lr=5e-4), torch.optim.Adam([b], lr=5e-4, betas=(0.9, 0.999))
    for i in range(5):
        gr = torch.randn(10007, generator=gen).to(DEV)
        a.grad, b.grad = gr.clone(), gr.clone()
        oa.step(), ob.step()
    assert rel_err(a, b) < 1e-6


def test_error_paths(nb, models):
    """Bad arguments raise (no silent fallback): CPU tensors, wrong channel count, coarse net asked for
    transient output."""
    c, f = models
    with pytest.raises(RuntimeError):
        nb.get_rays(H, W, FOCAL, torch.eye(4))
    with pytest.raises(RuntimeError):
        nb.raw2outputs_NeRFH_NFF(torch.zeros(2, 8, 100, device=DEV), torch.zeros(2, 8, device=DEV))
    with pytest.raises(RuntimeError):
        c.query(torch.zeros(2, 4, 3, device=DEV), torch.zeros(2, 3, device=DEV), 2)
    z = nb.raw2outputs_NeRFH_NFF(torch.zeros(0, 64, 132, device=DEV), torch.zeros(0, 64, device=DEV))
    assert z[0].shape == (0, 3)                                 # empty batch is a no-op, not an error


# ---------------------------------------------------------------------------------------------
# bf16 tensor-core path (tcgen05).  Stated tolerance: operands are rounded to bf16 (2^-9 relative) at every
# layer, accumulation is fp32; against the fp32 oracle the raw outputs agree to ~1e-2 of their scale.
# ---------------------------------------------------------------------------------------------
BF16_TOL = 3e-2


def nrm_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_bf16_mlp_forward_and_backward_vs_oracle(nb, weights, models):
    """Two references: (a) the fp32 oracle -- outputs within BF16_TOL; weight gradients within a loose bound,
    because a forward computed in bf16 flips the ReLU sign of pre-activations that are ~0 (a fraction p of
    flipped units moves a gradient by ~sqrt(p) in norm: inherent to any reduced-precision forward); (b) the
    oracle run with the SAME rounding points (bf16 operands, fp32 accumulate) -- gradients within 3e-2."""
    from nefes_b200 import _lib as L, ops
    wc, wf = weights
    c, f = models
    gen = torch.Generator().manual_seed(20)
    n, s = 37, 64                                     # 2368 points: 18.5 tiles -> exercises the ragged last tile
    pts = torch.rand(n, s, 3, generator=gen) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    for model, P, mode, typ, tr in ((f, wf, 2, "fine", True), (c, wc, 1, "coarse", False), (c, wc, 0, "coarse", False)):
        model.zero_grad()
        pg, dg = pts.to(DEV).requires_grad_(True), dirs.to(DEV).requires_grad_(True)
        raw = ops.field_query(pg, None if mode == 0 else dg, model.flat, model.net_id, mode, L.PREC_BF16)
        k = torch.randn(raw.shape, generator=gen)
        (raw * k.to(DEV)).sum().backward()
        views = model.layer_views(model.flat.grad)
        for emulate, tol_raw, tol_g in ((False, BF16_TOL, 0.2), (True, 4e-3, 3e-2)):
            Pg = O.clone_params(P, requires_grad=True)
            pr, dr = pts.clone().requires_grad_(True), dirs.clone().requires_grad_(True)
            ref = O.query_field(O.bf16_weights(Pg) if emulate else Pg, pr, dr, typ, tr, test_time=(mode == 0),
                                q=O.bf16_round if emulate else None)
            assert raw.shape == ref.shape
            assert rel_err(raw, ref) < tol_raw, (mode, emulate, rel_err(raw, ref))
            (ref * k).sum().backward()
            for key, ref_g in Pg.items():
                if ref_g.grad is None:
                    continue
                assert nrm_err(views[key], ref_g.grad) < tol_g, (mode, emulate, key, nrm_err(views[key], ref_g.grad))
            # gradients to the sample positions / view directions (pose refinement path)
            assert nrm_err(pg.grad, pr.grad) < 2 * tol_g, (mode, emulate, "d_pts", nrm_err(pg.grad, pr.grad))
            if mode != 0:
                assert nrm_err(dg.grad, dr.grad) < 2 * tol_g, (mode, emulate, "d_dirs", nrm_err(dg.grad, dr.grad))


def test_bf16_render_train_step(nb, weights, models):
    wc, wf = weights
    c, f = models
    gen = torch.Generator().manual_seed(21)
    n = 512
    o, d = O.camera_rays(H, W, FOCAL, torch.eye(4)[:3])
    pix = torch.randperm(H * W, generator=gen)[:n]
    rays = (o.reshape(-1, 3)[pix].contiguous(), d.reshape(-1, 3)[pix].contiguous())
    t_rand, u = torch.rand(n, 64, generator=gen), torch.rand(n, 64, generator=gen)
    with torch.no_grad():
        ref = O.render(H, W, FOCAL, wc, wf, rays=rays, near=NEAR, far=FAR, test_time=False, t_rand=t_rand, u=u)
    c.precision = f.precision = "bf16"
    try:
        c.zero_grad(), f.zero_grad()
        rgb, disp, acc, ex = nb.render(H, W, FOCAL, rays=(rays[0].to(DEV), rays[1].to(DEV)), img_idx=torch.zeros(1, 10),
                                       t_rand=t_rand.to(DEV), u=u.to(DEV), **render_kwargs(nb, models, False))
        for k, v in dict(rgb_map=rgb, feat_map=ex["feat_map"], rgb0=ex["rgb0"], acc_map=acc, beta=ex["beta"]).items():
            assert rel_err(v, ref[k]) < BF16_TOL, (k, rel_err(v, ref[k]))
        (rgb.mean() + ex["feat_map"].mean() + ex["rgb0"].mean()).backward()
        assert torch.isfinite(f.flat.grad).all() and torch.isfinite(c.flat.grad).all()
        assert float(f.flat.grad.abs().max()) > 0
    finally:
        c.precision = f.precision = "fp32"


def test_pose_refinement_matches_oracle_loop(nb, weights, models):
    """The pose-gradient CHAIN of the refinement on random-init fields (test_time render from c2w, cosine feature loss,
    so(3)+t delta): along the fp64 oracle's trajectory, at IDENTICAL poses, the engine's gradient of the 6 pose parameters
    is as close to the fp64 gradient as the fp32 reference's own gradient is (factor 3) -- measured ~5e-6 relative for
    rotation and ~3e-3 for translation on BOTH sides: with random-init fields the translation gradient is ~100x smaller
    and sums PE-backward terms scaled by up to 2^9, so it carries fp32 summation-order noise.

    The north-star bar on the REFINED POSE (1 mm / 0.01 deg after 50 iterations at 60x80) is asserted on a conditioned
    problem -- fields trained to carry pose signal -- in tests/test_gpu_refine_c4.py; on random-init fields Adam (which
    divides each step by |g|) turns the noise above into step-sized differences, for the fp32 reference as much as for
    the engine (it ends 1.3 mm from its own fp64 twin after 8 steps), so a free-running comparison here says nothing."""
    from nefes_b200 import refine
    wc, wf = weights
    c, f = models
    h, w_, focal = 30, 40, FOCAL / 2
    g = np_load_poses()
    gt = torch.tensor(g["test_gt"][0].reshape(3, 4), dtype=torch.float32)
    init = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32)
    with torch.no_grad():
        target = O.render(h, w_, focal, wc, wf, c2w=gt, near=NEAR, far=FAR, test_time=True)["feat_map"].t().contiguous()
    n_it, lr_r, lr_t = 8, 0.0087, 0.01
    for p in list(c.parameters()) + list(f.parameters()):
        p.requires_grad_(False)
    try:
        kw = render_kwargs(nb, models, True)
        Pc64, Pf64 = O.clone_params(wc, torch.float64), O.clone_params(wf, torch.float64)
        r64 = torch.zeros(3, dtype=torch.float64, requires_grad=True)
        t64 = torch.zeros(3, dtype=torch.float64, requires_grad=True)
        r32 = torch.zeros(3, requires_grad=True)
        t32 = torch.zeros(3, requires_grad=True)
        opt64 = torch.optim.Adam([{"params": [r64], "lr": lr_r}, {"params": [t64], "lr": lr_t}])
        opt32 = torch.optim.Adam([{"params": [r32], "lr": lr_r}, {"params": [t32], "lr": lr_t}])
        worst = {"r": (0., 0.), "t": (0., 0.)}
        for it in range(n_it):
            out = O.render(h, w_, focal, Pc64, Pf64, c2w=O.learn_pose_c2w(r64, t64, init.double()), near=NEAR, far=FAR,
                           test_time=True, hist=torch.zeros(1, 10, dtype=torch.float64))
            opt64.zero_grad()
            O.cosine_feature_loss(out["feat_map"].t(), target.double()).backward()
            # fp32 reference and engine at the SAME pose
            ra, ta = r64.detach().float().requires_grad_(True), t64.detach().float().requires_grad_(True)
            o32 = O.render(h, w_, focal, wc, wf, c2w=O.learn_pose_c2w(ra, ta, init), near=NEAR, far=FAR, test_time=True)
            O.cosine_feature_loss(o32["feat_map"].t(), target).backward()
            pm = refine.LearnPose(1, True, True, init[None].to(DEV)).to(DEV)
            with torch.no_grad():
                pm.r.copy_(r64.detach().float()[None]), pm.t.copy_(t64.detach().float()[None])
            rgb, disp, acc, ex = nb.render(h, w_, focal, c2w=pm(0)[:3, :4], img_idx=torch.zeros(1, 10), **kw)
            refine.feature_loss(ex["feat_map"].t(), target.to(DEV)).backward()
            for name, mine, a32, a64 in (("r", pm.r.grad[0].cpu(), ra.grad, r64.grad), ("t", pm.t.grad[0].cpu(), ta.grad, t64.grad)):
                scale = float(a64.abs().max())
                e_mine = float((mine.double() - a64).abs().max()) / scale
                e_ref = float((a32.double() - a64).abs().max()) / scale
                worst[name] = (max(worst[name][0], e_mine), max(worst[name][1], e_ref))
            opt64.step()
            # the reference's own free-running fp32 loop
            o = O.render(h, w_, focal, wc, wf, c2w=O.learn_pose_c2w(r32, t32, init), near=NEAR, far=FAR, test_time=True)
            opt32.zero_grad()
            O.cosine_feature_loss(o["feat_map"].t(), target).backward()
            opt32.step()
        print(f"worst gradient error vs fp64 over the trajectory (engine, fp32 reference): rotation {worst['r']}, translation {worst['t']}")
        assert worst["r"][0] <= max(3 * worst["r"][1], 1e-5) and worst["t"][0] <= max(3 * worst["t"][1], 1e-5), worst
        assert worst["r"][0] < 1e-4 and worst["t"][0] < 3e-2
        pose, losses = refine.refine_pose(init.to(DEV), target.to(DEV), h, w_, focal, kw, n_iters=n_it, lr_r=lr_r, lr_t=lr_t)
    finally:
        for p in list(c.parameters()) + list(f.parameters()):
            p.requires_grad_(True)
    ref32 = O.learn_pose_c2w(r32, t32, init).detach()
    dt, dang = O.pose_error(pose.cpu(), ref32)
    print(f"free-running 8-step loops on random-init fields, engine vs fp32 reference: {dt * 1e3:.3f} mm {dang:.5f} deg (informational)")
    assert dang < 1e-2, dang
    assert float(losses[-1]) < float(losses[0])


def np_load_poses():
    import numpy as np
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poses_stairs.npz"))
