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


def test_render_gradients_do_not_depend_on_netchunk():
    """netchunk splits the field query into several launches (TiledRaw.cat): the compact-cotangent shortcut must step
    aside there and the weight gradients must match the single-launch render."""
    import nefes_b200 as nb
    c, f = _nets()
    H, W, focal = 60, 80, 65.688
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 384
    ro = torch.zeros(n, 3, device="cuda")
    rd = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    t_rand, u = torch.rand(n, 64, device="cuda", generator=g), torch.rand(n, 64, device="cuda", generator=g)

    def run(netchunk):
        class Args:
            nerfh_nff, use_fine_only, NeRFW, transient_at_test = True, False, True, True
        Args.netchunk = netchunk
        q = lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(
            i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
        c.zero_grad(); f.zero_grad()
        rgb, disp, acc, ex = nb.render(H, W, focal, rays=(ro, rd), img_idx=torch.zeros(1, 10), near=0., far=4., ndc=False,
                                       use_viewdirs=True, network_query_fn=q, N_samples=64, N_importance=64, network_fn=c,
                                       network_fine=f, perturb=1., raw_noise_std=0., test_time=False, args=Args(),
                                       t_rand=t_rand, u=u)
        (rgb.sum() + (ex["feat_map"] ** 2).sum() + ex["rgb0"].sum() + ex["beta"].sum()).backward()
        return rgb.detach(), ex["feat_map"].detach(), c.flat.grad.clone(), f.flat.grad.clone()

    a, b = run(1 << 21), run(128 * 64)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for x, y in zip(a[2:], b[2:]):
        assert float((x - y).abs().max()) <= 1e-2 * float(x.abs().max())


def test_graph_replayed_refinement_matches_eager_loop():
    """PoseRefiner (one captured iteration replayed, cached across queries) against the eager refinement loop:
    same losses and the same refined pose, for two consecutive queries (the second one is replays only)."""
    import numpy as np, os
    import nefes_b200 as nb
    from nefes_b200 import refine
    c, f = _nets()
    # fp32 field arithmetic: with random-init weights the translation gradient is noise-level and Adam normalises it, so
    # in bf16 a 1-ulp difference in an update flips operand roundings and the two trajectories drift apart by
    # centimetres within a dozen steps; fp32 keeps the comparison about the graph mechanics
    c.precision = f.precision = "fp32"
    for p in (c.flat, f.flat):
        p.requires_grad_(False)
    H, W, focal = 60, 80, 65.688
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "poses_stairs.npz"))

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
    q = lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(
        i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
    kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True,
              white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=4., perturb=0., raw_noise_std=0.,
              test_time=True)
    gen = torch.Generator(device="cuda").manual_seed(11)
    for qi in range(2):
        init = torch.tensor(g["dfnet_init"][qi].reshape(3, 4), dtype=torch.float32, device="cuda")
        target = torch.randn(128, H * W, device="cuda", generator=gen)
        pe, le = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=False)
        pg, lg = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=True)
        # the loss trajectories must coincide; the poses only up to the loop's own run-to-run spread: the eager loop
        # itself is not bit-reproducible (atomic sums in the ray / encoding backward), and Adam amplifies that on the
        # noise-level translation gradient of random-init weights to ~5 mm over 12 steps (measured eager vs eager)
        for a, b in zip(le, lg):
            assert abs(float(a) - float(b)) < 2e-5, (float(a), float(b))
        assert float((pe[:, :3] - pg[:, :3]).abs().max()) < 2e-2
        assert float((pe[:, 3] - pg[:, 3]).abs().max()) < 6e-2


def test_large_ragged_render_is_chunk_invariant():
    """C5-style inference render: 40 000 rays (not a multiple of the 32 768-ray chunk nor of the tile), test_time, no
    gradients; the result must not depend on how batchify_rays cuts it."""
    import nefes_b200 as nb
    c, f = _nets()
    H, W, focal = 60, 106, 93.0
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 40000 + 37
    ro = torch.randn(n, 3, device="cuda", generator=g) * 0.1
    rd = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
    q = lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(
        i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
    kw = dict(rays=(ro, rd), img_idx=torch.zeros(1, 10), near=0., far=10., ndc=False, use_viewdirs=True, network_query_fn=q,
              N_samples=64, N_importance=64, network_fn=c, network_fine=f, perturb=0., raw_noise_std=0., test_time=True,
              args=Args())
    with torch.no_grad():
        a = nb.render(H, W, focal, chunk=32768, **kw)
        b = nb.render(H, W, focal, chunk=8192 + 64, **kw)
    assert a[0].shape == (n, 3) and a[3]["feat_map"].shape == (n, 128)
    assert torch.isfinite(a[0]).all() and torch.isfinite(a[3]["feat_map"]).all()
    assert torch.equal(a[0], b[0]) and torch.equal(a[3]["feat_map"], b[3]["feat_map"]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("precision,test_time", [("bf16", False), ("bf16", True), ("fp32", False)])
def test_one_call_render_rays_matches_staged_path(precision, test_time):
    """nefes_render_rays_fwd/_bwd (render_rays as ONE engine call, reached through nb.StandardQuery) against the staged
    path (a hand-written query closure: one engine call per stage) on the same draws: the same kernels run on the same
    bits, so outputs are identical; gradients differ only by the order of fp32 atomics / reductions."""
    import nefes_b200 as nb
    c, f = _nets()
    c.precision = f.precision = precision
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 333                                                   # ragged: last tile partly filled
    t_rand, u = torch.rand(n, 64, device="cuda", generator=g), torch.rand(n, 64, device="cuda", generator=g)
    ro0 = torch.randn(n, 3, device="cuda", generator=g) * 0.1
    rd0 = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1) * 1.3

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21

    def run(q):
        c.zero_grad(); f.zero_grad()
        ro, rd = ro0.clone().requires_grad_(True), rd0.clone().requires_grad_(True)
        rgb, disp, acc, ex = nb.render(60, 80, 65.688, rays=(ro, rd), img_idx=torch.zeros(1, 10), near=0., far=4., ndc=False,
                                       use_viewdirs=True, network_query_fn=q, N_samples=64, N_importance=64, network_fn=c,
                                       network_fine=f, perturb=0. if test_time else 1., raw_noise_std=0., test_time=test_time,
                                       args=Args(), t_rand=t_rand, u=u, return_aux=True)
        loss = rgb.sum() + (ex["feat_map"] ** 2).sum() + disp.mean() + acc.sum()
        if not test_time:
            loss = loss + ex["rgb0"].sum() + ex["beta"].sum() + ex["feat0"].abs().sum() + 0.1 * ex["transient_sigmas"].mean()
        loss.backward()
        out = dict(rgb=rgb, disp=disp, acc=acc, **ex)
        grads = dict(ro=ro.grad, rd=rd.grad)
        if not test_time:
            grads.update(c=c.flat.grad.clone(), f=f.flat.grad.clone())
        return {k: v.detach() for k, v in out.items()}, grads

    staged = run(lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(
        i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb))
    fused = run(nb.StandardQuery(Args.netchunk))
    assert set(staged[0]) == set(fused[0])
    for k, v in staged[0].items():
        if k == "z_std":
            assert float((v - fused[0][k]).abs().max()) < 1e-6
        else:
            assert torch.equal(v, fused[0][k]), k
    for k, v in staged[1].items():
        assert float((v - fused[1][k]).abs().max()) <= 2e-3 * float(v.abs().max()) + 1e-7, k
    if test_time:                                           # frozen-field refinement: no parameter gradient is produced
        assert c.flat.grad is None or float(c.flat.grad.abs().max()) == 0.0


WORLD = {"pose_scale": 0.75, "move_all_cam_vec": [0.1, -0.3, 0.2], "pose_scale2": 1.6}


@pytest.mark.parametrize("lietorch,world", [(False, None), (True, None), (True, WORLD), (False, WORLD)])
def test_refinement_glue_kernels_match_torch(lietorch, world):
    """nefes_pose_rays_fwd/_bwd, nefes_cosine_loss_fwd/_bwd and nefes_pose_adam_step against the torch chain they
    replace: LearnPose (so(3) exponential, or SE3.exp([t, r]) with lietorch=True) -> fix_coord_supp -> get_rays ->
    render()'s packing; F.cosine_similarity loss; autograd to (r, t); torch.optim.Adam with two parameter groups -- three
    consecutive steps from a non-zero rotation."""
    import nefes_b200 as nb
    from nefes_b200 import _lib as L, refine
    lib, p = L.lib(), L.ptr
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    H, W, focal, C_ = 12, 20, 21.5, 64
    N = H * W
    g = torch.Generator(device="cuda").manual_seed(21)
    init = torch.linalg.qr(torch.randn(3, 3, device=dev, generator=g))[0]
    init = torch.cat([init, torch.randn(3, 1, device=dev, generator=g)], 1).contiguous()
    target = torch.randn(C_, N, device=dev, generator=g)
    mix = torch.randn(21, C_, device=dev, generator=g)            # a differentiable stand-in for the render: feat = rays @ mix
    pose_t = refine.LearnPose(1, True, True, init[None], lietorch=lietorch).to(dev)
    chain = refine._chain6(lietorch, world)
    with torch.no_grad():
        pose_t.r.copy_(torch.tensor([[0.03, -0.02, 0.05]]))
        pose_t.t.copy_(torch.tensor([[0.1, 0.0, -0.2]]))
    opt = torch.optim.Adam([{"params": [pose_t.r], "lr": 0.0087}, {"params": [pose_t.t], "lr": 0.01}])
    pose6 = torch.cat([pose_t.r.detach().reshape(-1), pose_t.t.detach().reshape(-1)]).contiguous()
    rays, c2w, d_c2w = torch.empty(N, 21, device=dev), torch.empty(3, 4, device=dev), torch.zeros(12, device=dev)
    stats, state, loss, hist = torch.zeros(3, C_, device=dev), torch.zeros(13, device=dev), torch.zeros(1, device=dev), torch.zeros(8, device=dev)
    d_feat = torch.empty(N, C_, device=dev)
    for it in range(3):
        # torch chain
        m = pose_t(0)
        if world is not None:
            m = refine.fix_coord_supp(None, m[None, :3, :4], world)[0]
        ro, rd = nb.get_rays(H, W, focal, m[:3, :4])
        rd_f, ro_f = rd.reshape(-1, 3), ro.reshape(-1, 3)
        vd = rd_f / torch.norm(rd_f, dim=-1, keepdim=True)
        rb = torch.cat([ro_f, rd_f, torch.full((N, 1), 0.5, device=dev), torch.full((N, 1), 4.0, device=dev), vd,
                        torch.zeros(N, 10, device=dev)], 1)
        feat_t = rb @ mix
        loss_t = refine.feature_loss(feat_t.t(), target)
        opt.zero_grad()
        loss_t.backward()
        # engine chain
        L.check(lib.nefes_pose_rays_fwd(p(pose6), p(init), H, W, focal, 0.5, 4.0, p(c2w), p(rays), 21, chain, st), "fwd")
        assert float((c2w - m[:3, :4]).abs().max()) < 2e-6
        assert float((rays - rb).abs().max()) < 1e-5
        feat = (rays @ mix).contiguous()
        L.check(lib.nefes_cosine_loss_fwd(p(feat), p(target), None, N, C_, p(stats), st), "loss fwd")
        L.check(lib.nefes_cosine_loss_bwd(p(feat), p(target), None, p(stats), N, C_, p(loss), p(hist), p(state[12:]), 8, p(d_feat), st), "loss bwd")
        assert abs(float(loss) - float(loss_t)) < 1e-6 and abs(float(hist[it]) - float(loss_t)) < 1e-6
        d_rays = (d_feat @ mix.t()).contiguous()
        L.check(lib.nefes_pose_rays_bwd(p(d_rays), p(rays), 21, H, W, focal, p(d_c2w), st), "bwd")
        opt.step()
        L.check(lib.nefes_pose_adam_step(p(pose6), p(init), p(d_c2w), p(stats), 3 * C_, p(state), 0.0087, 0.01, 0.9, 0.999, 1e-8, chain, st), "adam")
        assert float(d_c2w.abs().max()) == 0.0 and float(stats.abs().max()) == 0.0 and float(state[12]) == it + 1
        ref6 = torch.cat([pose_t.r.detach().reshape(-1), pose_t.t.detach().reshape(-1)])
        # Adam's first steps are lr * sign(g): any error in the gradient chain shows up at full step size
        assert float((pose6 - ref6).abs().max()) < 2e-5, (it, pose6, ref6)


def test_engine_refinement_iteration_matches_torch_loop():
    """EnginePoseRefiner (the iteration as ~30 engine launches, replayed from a CUDA graph) against the loop with the torch
    pose chain / loss / optimiser around the engine render: same loss trajectory, same refined pose up to the loop's own
    run-to-run spread (see test_graph_replayed_refinement_matches_eager_loop), for two consecutive queries."""
    import numpy as np, os
    import nefes_b200 as nb
    from nefes_b200 import refine
    c, f = _nets()
    c.precision = f.precision = "fp32"
    for p in (c.flat, f.flat):
        p.requires_grad_(False)
    H, W, focal = 60, 80, 65.688
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "poses_stairs.npz"))

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
    kw = dict(network_query_fn=nb.StandardQuery(Args.netchunk), N_importance=64, N_samples=64, network_fn=c, network_fine=f,
              use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=4., perturb=0.,
              raw_noise_std=0., test_time=True)
    gen = torch.Generator(device="cuda").manual_seed(11)
    for qi in range(2):
        init = torch.tensor(g["dfnet_init"][qi].reshape(3, 4), dtype=torch.float32, device="cuda")
        target = torch.randn(128, H * W, device="cuda", generator=gen)
        pe, le = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=False, engine=False)
        pg, lg = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=True, engine=True)
        assert len(le) == len(lg) == 12
        for a, b in zip(le, lg):
            assert abs(float(a) - float(b)) < 2e-5, (float(a), float(b))
        assert float((pe[:, :3] - pg[:, :3]).abs().max()) < 2e-2
        assert float((pe[:, 3] - pg[:, 3]).abs().max()) < 6e-2
    # bf16 fields go the same way (trajectories drift with operand roundings; the first losses must still agree)
    c.precision = f.precision = "bf16"
    pe, le = refine.refine_pose(init, target, H, W, focal, kw, n_iters=4, graph=False, engine=False)
    pg, lg = refine.refine_pose(init, target, H, W, focal, kw, n_iters=4, graph=False, engine=True)
    assert abs(float(le[0]) - float(lg[0])) < 1e-5


def test_one_call_render_edge_cases():
    """Empty ray batch; ray chunks (`chunk` < N: batchify_rays cuts the batch, every chunk is one engine call) against the
    single call; a netchunk smaller than the batch sends render_rays down the staged route with the same numbers."""
    import nefes_b200 as nb
    c, f = _nets()
    g = torch.Generator(device="cuda").manual_seed(13)
    n = 700
    ro = torch.randn(n, 3, device="cuda", generator=g) * 0.1
    rd = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    t_rand, u = torch.rand(n, 64, device="cuda", generator=g), torch.rand(n, 64, device="cuda", generator=g)

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test = True, False, True, True

    def run(rays, chunk, netchunk, tr, uu):
        kw = dict(rays=rays, img_idx=torch.zeros(1, 10), near=0., far=4., ndc=False, use_viewdirs=True,
                  network_query_fn=nb.StandardQuery(netchunk), N_samples=64, N_importance=64, network_fn=c, network_fine=f,
                  perturb=1., raw_noise_std=0., test_time=False, args=Args(), t_rand=tr, u=uu)
        with torch.no_grad():
            rgb, disp, acc, ex = nb.render(60, 80, 65.688, chunk=chunk, **kw)
        return rgb, disp, acc, ex

    one = run((ro, rd), 32768, 1 << 21, t_rand, u)
    cut = run((ro, rd), 256, 1 << 21, t_rand, u)                 # 3 ray chunks, each one engine call
    staged = run((ro, rd), 32768, 128 * 64, t_rand, u)           # netchunk < batch: staged route, 128-ray field queries
    for other in (cut, staged):
        for a, b in zip(one[:3], other[:3]):
            assert torch.equal(a, b)
        for k in ("feat_map", "rgb0", "beta", "transient_sigmas"):
            assert torch.equal(one[3][k], other[3][k]), k
        assert float((one[3]["z_std"] - other[3]["z_std"]).abs().max()) < 1e-6
    empty = run((ro[:0], rd[:0]), 32768, 1 << 21, t_rand[:0], u[:0])
    assert empty[0].shape == (0, 3) and empty[3]["feat_map"].shape == (0, 128) and empty[3]["transient_sigmas"].shape == (0, 128)


def test_gather_ray_batch_matches_the_reference_loops():
    """batching.gather_ray_batch against run_nefes.py:66-74 written as the reference writes it (per-image fancy
    indexing + cat), for per-image selections and for the shared patch selection."""
    import nefes_b200 as nb
    g = torch.Generator(device="cuda").manual_seed(17)
    B, H, W, n = 3, 12, 20, 40
    pose = torch.cat([torch.linalg.qr(torch.randn(B, 3, 3, device="cuda", generator=g))[0],
                      torch.randn(B, 3, 1, device="cuda", generator=g)], -1)
    target = torch.rand(B, H, W, 3, device="cuda", generator=g)
    ftarget = torch.randn(B, H, W, 16, device="cuda", generator=g)
    hist = torch.rand(B, 10, device="cuda", generator=g)
    ro, rd = nb.get_rays_batch(H, W, 21.5, pose)
    for sel in (nb.select_random_pixels(B, H, W, n, device="cuda", generator=g),
                nb.select_random_patches(H, W, num_crops=2, crop_size=4, device="cuda", generator=g)):
        rays, ts, tf, he = nb.gather_ray_batch(H, W, 21.5, pose, sel, target, ftarget, hist)
        selb = sel if sel.dim() == 2 else sel[None].expand(B, -1)
        rows, cols = selb // W, selb % W
        ref_o = torch.cat([ro[k, rows[k], cols[k]] for k in range(B)])
        ref_d = torch.cat([rd[k, rows[k], cols[k]] for k in range(B)])
        ref_t = torch.cat([target[k, rows[k], cols[k]] for k in range(B)])
        ref_f = torch.cat([ftarget[k, rows[k], cols[k]] for k in range(B)])
        ref_h = torch.cat([hist[k].expand(selb.shape[1], 10) for k in range(B)])
        assert torch.equal(rays[0], ref_o) and torch.equal(rays[1], ref_d)
        assert torch.equal(ts, ref_t) and torch.equal(tf, ref_f) and torch.equal(he, ref_h)


def test_plain_c_host_renders_through_the_c_abi(tmp_path):
    """The C99 example host (no Python, no torch in the process) runs render_rays forward + backward through
    libnefes_b200.so and gets finite outputs and non-zero weight gradients for both fields."""
    import subprocess
    from test_abi import _build_c_example
    exe = _build_c_example(tmp_path)
    if exe is None:
        pytest.skip("gcc or the CUDA runtime headers are not available")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok") and "|d params fine|" in r.stdout


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mlp_dgrad_and_wgrad_halves_match_the_combined_backward(precision):
    """nefes_mlp_dgrad + nefes_mlp_wgrad (SURVEY 8b's export names) against nefes_mlp_bwd on the same saved state."""
    import ctypes as C
    from nefes_b200 import _lib as L
    from nefes_b200 import ops
    c, f = _nets()
    lib, p = L.lib(), L.ptr
    prec = L.PREC_BF16 if precision == "bf16" else L.PREC_FP32
    g = torch.Generator(device="cuda").manual_seed(31)
    N, S, mode, net = 70, 64, L.MODE_FULL, L.NET_FINE
    pts = torch.rand(N, S, 3, device="cuda", generator=g) * 2 - 1
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=g), dim=-1)
    flat = f.flat.detach()
    sv, sf, sb = L.mlp_workspace(net, mode, prec, N * S, N)
    saved, scr_f, scr_b = ops._buf(sv, "cuda"), ops._buf(sf, "cuda"), ops._buf(sb, "cuda")
    raw = torch.empty(N, S, 137, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.nefes_mlp_fwd(p(flat), net, mode, prec, p(pts), p(dirs), N, S, p(raw), p(saved), p(scr_f), st), "fwd")
    d_raw = torch.randn(N, S, 137, device="cuda", generator=g)
    dP, dp, dd = torch.zeros_like(flat), torch.empty_like(pts), torch.empty_like(dirs)
    L.check(lib.nefes_mlp_bwd(p(flat), net, mode, prec, p(pts), p(dirs), N, S, p(raw), p(d_raw), p(saved), p(scr_b), p(dP), p(dp),
                              p(dd), st), "bwd")
    dp2, dd2, dP2 = torch.empty_like(pts), torch.empty_like(dirs), torch.zeros_like(flat)
    L.check(lib.nefes_mlp_dgrad(p(flat), net, mode, prec, p(pts), p(dirs), N, S, p(raw), p(d_raw), p(saved), p(scr_b), p(dp2), p(dd2),
                                st), "dgrad")
    L.check(lib.nefes_mlp_wgrad(p(flat), net, mode, prec, p(pts), p(dirs), N, S, p(raw), p(d_raw), p(saved), p(scr_b), p(dP2), st),
            "wgrad")
    tol = 1e-5 if precision == "fp32" else 2e-2        # bf16: the combined call takes the fused wgrad launches (other rounding points)
    for a, b in ((dp, dp2), (dd, dd2), (dP, dP2)):
        assert float((a - b).abs().max()) <= tol * float(a.abs().max()) + 1e-8
    assert lib.nefes_mlp_dgrad(p(flat), net, mode, prec, p(pts), p(dirs), N, S, p(raw), p(d_raw), p(saved), p(scr_b), None, None, st) == 1
    assert lib.nefes_mlp_wgrad(p(flat), net, mode, prec, p(pts), p(dirs), N, S, p(raw), p(d_raw), p(saved), p(scr_b), None, st) == 1


def test_render_beyond_one_netchunk_takes_the_one_call_route_in_groups():
    """The reference's defaults (chunk 32768 rays, netchunk 2^21 points) put two netchunks of fine points into one render_rays
    call: the one-call route runs per netchunk-sized group of rays and must give the numbers (and gradients) of smaller calls."""
    import nefes_b200 as nb
    dev = torch.device("cuda")

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 14
    c = nb.NeRFH_NFF("coarse", W=128, precision="bf16").to(dev)
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True, precision="bf16").to(dev)
    g = torch.Generator(device="cuda").manual_seed(2)
    n = 300                                             # 300 rays x 128 fine points = 2.3 netchunks of 2^14 points
    ro = torch.randn(n, 3, device=dev, generator=g) * 0.2
    rd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1)
    t_rand, u = torch.rand(n, 64, device=dev, generator=g), torch.rand(n, 64, device=dev, generator=g)

    def run(netchunk, chunk):
        kw = dict(network_query_fn=nb.StandardQuery(netchunk), N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True,
                  white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=4., perturb=1., raw_noise_std=0., test_time=False,
                  t_rand=t_rand, u=u)
        c.zero_grad(), f.zero_grad()
        rgb, disp, acc, ex = nb.render(60, 80, 65.7, chunk=chunk, rays=(ro, rd), img_idx=torch.zeros(1, 10), **kw)
        (rgb.sum() + ex["feat_map"].mean() + ex["rgb0"].sum() + ex["beta"].sum()).backward()
        return dict(rgb=rgb, disp=disp, acc=acc, **ex), c.flat.grad.clone(), f.flat.grad.clone()
    a, gca, gfa = run(1 << 14, 32768)                   # one render_rays call, three engine calls
    b, gcb, gfb = run(1 << 21, 100)                     # three render_rays calls of 100 rays, one engine call each
    for k in a:
        assert torch.equal(a[k], b[k]) or float((a[k] - b[k]).abs().max()) < 1e-6 * float(b[k].abs().max()), k
    assert float((gca - gcb).abs().max()) < 2e-3 * float(gcb.abs().max()) and float((gfa - gfb).abs().max()) < 2e-3 * float(gfb.abs().max())


def test_render_call_with_frozen_weights_packs_once():
    """nefes_render_rays_prepack + cfg.weights_packed (the refiner packs the frozen fields once per query): the forward that
    skips its re-pack is bit-equal to the one that re-packs; it keeps rendering the packed weights after the parameters
    change (the documented contract) until prepack() runs again."""
    from nefes_b200 import ops, _lib as L
    from nefes_b200.nerfh_nff import _PREC
    c, f = _nets()
    c.precision = f.precision = "bf16"
    n = 333
    cfg = dict(n_samples=64, n_importance=64, prec=_PREC["bf16"], test_time=True, output_transient=True, transient_at_test=True,
               net_coarse=c.net_id, net_fine=f.net_id, beta_min=f.beta_min)
    gen = torch.Generator(device="cuda").manual_seed(3)
    rays = torch.zeros(n, 21, device="cuda")
    rays[:, 0:3] = torch.randn(n, 3, device="cuda", generator=gen) * 0.1
    d = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=gen), dim=-1)
    rays[:, 3:6], rays[:, 8:11], rays[:, 7] = d, d, 4.0
    plain = ops.RenderCall(n, 21, cfg, c.flat, f.flat, "cuda")
    frozen = ops.RenderCall(n, 21, cfg, c.flat, f.flat, "cuda", frozen_weights=True)
    for rc in (plain, frozen):
        rc.rays.copy_(rays)
    plain.forward()
    frozen.prepack()
    frozen.forward()
    torch.cuda.synchronize()
    assert torch.equal(plain.feat, frozen.feat) and torch.equal(plain.rgb, frozen.rgb) and torch.equal(plain.w, frozen.w)
    feat0 = frozen.feat.clone()
    with torch.no_grad():
        f.flat.mul_(1.05)
    frozen.forward()                                     # still the packed images
    assert torch.equal(frozen.feat, feat0)
    frozen.prepack()
    frozen.forward()
    plain.forward()
    torch.cuda.synchronize()
    assert not torch.equal(frozen.feat, feat0) and torch.equal(plain.feat, frozen.feat)
    with torch.no_grad():
        f.flat.div_(1.05)
