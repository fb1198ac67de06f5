"""CPU, build container: the conditioned C4 refinement problem (VERDICT r1 "next" 3) on the oracle.
TEST INFRASTRUCTURE.  Inputs: tests/golden/c4_fields.npz (fields trained by tools/make_conditioned_fields.py so that the
rendered features depend on the pose).  Problem: ground truth = stairs test pose 0, target = the oracle's fp32 feature map at
the ground truth, start = DFNet's prediction for that image (0.29 m / 4.1 deg off), 50 Adam iterations at the full 60x80
render (lr_r 0.0087, lr_t 0.01, cosine feature loss, pose delta through so(3) x R^3 -- poses.py:25-50 with lietorch=False,
the reference's own pure-torch chain, whose arithmetic the oracle is pinned to).  Runs the loop on the oracle in fp32 and in
fp64 and writes tests/golden/c4_refine.npz: both trajectories, both loss curves, a subsample of the target for pinning.
The fp32-vs-fp64 spread of the oracle is the noise floor of this problem: an engine cannot be asked to be closer to the fp32
reference than the fp32 reference is to exact arithmetic.

    python oracle/make_c4_fixture.py [n_iters]        (about 10 minutes on 8 cores)
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import nefes_oracle as O      # noqa: E402

H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.
LR_R, LR_T = 0.0087, 0.01


def load_fields(dtype=torch.float32):
    z = np.load(os.path.join(ROOT, "tests", "golden", "c4_fields.npz"))
    Pc = {k[len("coarse/"):]: torch.from_numpy(z[k]).to(dtype) for k in z.files if k.startswith("coarse/")}
    Pf = {k[len("fine/"):]: torch.from_numpy(z[k]).to(dtype) for k in z.files if k.startswith("fine/")}
    return Pc, Pf


def problem():
    g = np.load(os.path.join(ROOT, "tests", "golden", "poses_stairs.npz"))
    gt = torch.tensor(g["test_gt"][0].reshape(3, 4), dtype=torch.float32)
    init = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32)
    return gt, init


def target_features(Pc, Pf, gt):
    with torch.no_grad():
        return O.render(H, W, FOCAL, Pc, Pf, c2w=gt, near=NEAR, far=FAR, test_time=True)["feat_map"].t().contiguous()


def run_loop(Pc, Pf, init, target, n_iters, dtype):
    r = torch.zeros(3, dtype=dtype, requires_grad=True)
    t = torch.zeros(3, dtype=dtype, requires_grad=True)
    opt = torch.optim.Adam([{"params": [r], "lr": LR_R}, {"params": [t], "lr": LR_T}])
    hist = torch.zeros(1, 10, dtype=dtype)
    poses, losses = [], []
    for it in range(n_iters):
        t0 = time.perf_counter()
        out = O.render(H, W, FOCAL, Pc, Pf, c2w=O.learn_pose_c2w(r, t, init.to(dtype)), near=NEAR, far=FAR, test_time=True, hist=hist)
        loss = O.cosine_feature_loss(out["feat_map"].t(), target.to(dtype))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
        poses.append(O.learn_pose_c2w(r, t, init.to(dtype)).detach().double().numpy().copy())
        print(f"  [{dtype}] iter {it}: loss {float(loss):.6f}  ({time.perf_counter() - t0:.1f} s)", flush=True)
    return np.stack(poses), np.asarray(losses)


def main():
    n_iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    torch.set_num_threads(os.cpu_count() or 8)
    Pc, Pf = load_fields()
    gt, init = problem()
    target = target_features(Pc, Pf, gt)
    print(f"target feature map: std over pixels {float(target.std(1).mean()):.4f} (random-init fields: ~0.02)")
    p32, l32 = run_loop(Pc, Pf, init, target, n_iters, torch.float32)
    Pc64, Pf64 = load_fields(torch.float64)
    p64, l64 = run_loop(Pc64, Pf64, init, target, n_iters, torch.float64)
    e0 = O.pose_error(init, gt)
    e32 = O.pose_error(torch.from_numpy(p32[-1]).float(), gt)
    d = O.pose_error(torch.from_numpy(p32[-1]), torch.from_numpy(p64[-1]))
    print(f"start {e0[0] * 1e3:.1f} mm / {e0[1]:.3f} deg from the ground truth; after {n_iters} iterations (fp32 oracle) {e32[0] * 1e3:.1f} mm / {e32[1]:.3f} deg; "
          f"fp32 oracle vs fp64 oracle: {d[0] * 1e3:.4f} mm / {d[1]:.5f} deg")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c4_refine.npz"), poses32=p32, poses64=p64, loss32=l32, loss64=l64,
                        target_sub=target[:, ::50].numpy(), n_iters=n_iters, lr=np.asarray([LR_R, LR_T]), gt=gt.numpy(), init=init.numpy())


if __name__ == "__main__":
    main()
