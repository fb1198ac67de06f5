"""GPU: the conditioned C4 refinement problem (BASELINE configs[3]): 50 pose-gradient iterations at the full
60x80 render on fields TRAINED to carry pose signal (tests/golden/c4_fields.npz, tools/make_conditioned_fields.py), the
engine-resident iteration in fp32 and in bf16 against the oracle's fp32 loop (tests/golden/c4_refine.npz, produced by
oracle/make_c4_fixture.py in the build container; the oracle's own fp64 loop gives the noise floor of the problem).

North-star bar: refined pose within 1 mm / 0.01 deg of the reference's."""
import os

import numpy as np
import pytest
import torch

from oracle import nefes_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21


@pytest.fixture(scope="module")
def problem():
    z = np.load(os.path.join(GOLD, "c4_fields.npz"))
    wc = {k[len("coarse/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("coarse/")}
    wf = {k[len("fine/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("fine/")}
    fx = {k: v for k, v in np.load(os.path.join(GOLD, "c4_refine.npz")).items()}
    gt, init = torch.from_numpy(fx["gt"]), torch.from_numpy(fx["init"])
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():                              # the oracle's fp32 render at the ground truth: the query's target
        target = O.render(H, W, FOCAL, wc, wf, c2w=gt, near=NEAR, far=FAR, test_time=True)["feat_map"].t().contiguous()
    # pinned: the same target the fixture's trajectories were computed against (thread-count dependent summation order only)
    assert float((target[:, ::50] - torch.from_numpy(fx["target_sub"])).abs().max()) < 2e-5
    return wc, wf, fx, init, target


def engine_kwargs(nb, wc, wf, prec):
    c = nb.NeRFH_NFF("coarse", W=128, precision=prec)
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True, precision=prec)
    c.load_state_dict(wc, strict=False)
    f.load_state_dict(wf)
    c, f = c.to(DEV), f.to(DEV)
    for p in list(c.parameters()) + list(f.parameters()):
        p.requires_grad_(False)
    return dict(network_query_fn=nb.StandardQuery(Args.netchunk), N_importance=64, N_samples=64, network_fn=c, network_fine=f,
                use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR, perturb=0.,
                raw_noise_std=0., test_time=True)


# the bars: (translation mm, rotation deg) of the final pose against the fp32 oracle's final pose
BARS = {"fp32": (1.0, 0.01), "bf16": (1.0, 0.01)}


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_c4_refined_pose_vs_oracle(problem, prec):
    import nefes_b200 as nb
    from nefes_b200 import refine
    wc, wf, fx, init, target = problem
    n_it = int(fx["n_iters"])
    kw = engine_kwargs(nb, wc, wf, prec)
    refine.clear_refiner_cache()
    pose, losses = refine.refine_pose(init.to(DEV), target.to(DEV), H, W, FOCAL, kw, n_iters=n_it, lr_r=float(fx["lr"][0]),
                                      lr_t=float(fx["lr"][1]))
    ref32 = torch.from_numpy(fx["poses32"][-1])
    ref64 = torch.from_numpy(fx["poses64"][-1])
    gt = torch.from_numpy(fx["gt"])
    dt, dang = O.pose_error(pose.cpu().double(), ref32)
    st, sang = O.pose_error(ref32, ref64)                                # the oracle's own fp32-vs-fp64 spread on this problem
    moved = O.pose_error(ref32, init.double())
    left = O.pose_error(ref32, gt.double())
    mine_left = O.pose_error(pose.cpu().double(), gt.double())
    l_eng = np.asarray([float(x) for x in losses])
    print(f"[{prec}] engine vs fp32 oracle after {n_it} iterations: {dt * 1e3:.3f} mm / {dang:.5f} deg   "
          f"(oracle fp32 vs fp64: {st * 1e3:.3f} mm / {sang:.5f} deg; pose travelled {moved[0] * 1e3:.1f} mm / {moved[1]:.2f} deg; "
          f"distance left to the ground truth: oracle {left[0] * 1e3:.1f} mm / {left[1]:.3f} deg, engine {mine_left[0] * 1e3:.1f} mm / {mine_left[1]:.3f} deg); "
          f"loss curve max |diff| {np.abs(l_eng - fx['loss32']).max():.2e}")
    assert moved[0] > 0.05 and moved[1] > 1.0, "the refinement did not move the pose: test is vacuous"
    assert l_eng[-1] < 0.5 * l_eng[0]
    bar_t, bar_r = BARS[prec]
    if prec == "fp32":
        # 1 mm / 0.01 deg, or twice the reference's own distance from exact arithmetic where that is larger
        assert dt * 1e3 <= max(bar_t, 2 * st * 1e3), (dt, st)
        assert dang <= max(bar_r, 2 * sang), (dang, sang)
        assert np.abs(l_eng - fx["loss32"]).max() < 1e-4
    else:
        # bf16 fields (the path bench.py times): the SAME 1 mm / 0.01 deg bar (measured 0.48 mm / 0.000 deg)
        assert dt * 1e3 <= bar_t and dang <= bar_r, (dt, dang)
        assert np.abs(l_eng - fx["loss32"]).max() < 2e-4
        # and the refinement does its job as well as the fp32 reference does: no further from the ground truth
        assert mine_left[0] <= left[0] + 1e-3 and mine_left[1] <= left[1] + 0.01


def test_refinement_with_fusion_stage_vs_oracle(problem):
    """The reference's full refinement step on nerfh_nff configs (DFM_pose_refine.py:321-337): render -> FusionNet(rgb, feat_map)
    -> cosine loss on the FUSED features, so the pose gradient also flows through the rendered colours.  Three Adam iterations
    at 30x40 on the engine (render, FusionNet and their backwards as engine launches, iteration replayed from a CUDA graph)
    against the same loop on the oracle (fp32, CPU): Adam's first steps are lr * sign(g), any error in the chain shows at
    full step size."""
    import nefes_b200 as nb
    import nefes_b200.nerfh_nff as NB
    from nefes_b200 import refine
    wc, wf, fx, init, _ = problem
    h, w_, focal = 30, 40, FOCAL / 2
    torch.manual_seed(13)
    fus = NB.FusionNet(128).eval()
    Pfus = {k: v.clone() for k, v in fus.state_dict().items()}
    gt = torch.from_numpy(fx["gt"])
    with torch.no_grad():
        o = O.render(h, w_, focal, wc, wf, c2w=gt, near=NEAR, far=FAR, test_time=True)
        target = O.fusion_net(Pfus, o["rgb_map"], o["feat_map"], 1, h, w_, training=False)[0].reshape(128, -1).contiguous()
    # oracle loop
    r, t = torch.zeros(3, requires_grad=True), torch.zeros(3, requires_grad=True)
    opt = torch.optim.Adam([{"params": [r], "lr": 0.0087}, {"params": [t], "lr": 0.01}])
    ref_losses = []
    for _ in range(3):
        o = O.render(h, w_, focal, wc, wf, c2w=O.learn_pose_c2w(r, t, init), near=NEAR, far=FAR, test_time=True)
        f = O.fusion_net(Pfus, o["rgb_map"], o["feat_map"], 1, h, w_, training=False)[0].reshape(128, -1)
        loss = O.cosine_feature_loss(f, target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ref_losses.append(float(loss))
    ref_pose = O.learn_pose_c2w(r, t, init).detach()
    kw = engine_kwargs(nb, wc, wf, "fp32")
    kw["network_fn"].fusion_net = fus.to(DEV)
    for p in kw["network_fn"].parameters():
        p.requires_grad_(False)
    for graph in (False, True):
        refine.clear_refiner_cache()
        pose, losses = refine.refine_pose(init.to(DEV), target.to(DEV), h, w_, focal, kw, n_iters=3, fusion=True, graph=graph)
        dt, dang = O.pose_error(pose.cpu().double(), ref_pose.double())
        print(f"[fusion, graph={graph}] engine vs oracle after 3 iterations: {dt * 1e3:.4f} mm / {dang:.5f} deg; losses {[float(x) for x in losses]} vs {ref_losses}")
        assert dt < 1e-3 and dang < 1e-2
        assert max(abs(float(a) - b) for a, b in zip(losses, ref_losses)) < 1e-5
