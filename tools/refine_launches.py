"""Eager refinement iterations (C4 shape, bf16) for an ncu launch list: which kernels one iteration launches.
Usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/refine_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nefes_b200 as nb
from nefes_b200 import refine

H, W, FOCAL = 60, 80, 65.688
dev = torch.device("cuda")
c = nb.NeRFH_NFF("coarse", W=128).to(dev)
f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True).to(dev)
c.precision = f.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for p in (c.flat, f.flat):
    p.requires_grad_(False)


class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21


kw = dict(network_query_fn=nb.StandardQuery(Args.netchunk), N_importance=64, N_samples=64, network_fn=c, network_fine=f,
          use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=4., perturb=0.,
          raw_noise_std=0., test_time=True)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "poses_stairs.npz"))
init = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32, device=dev)
target = torch.randn(128, H * W, device=dev)
mode = sys.argv[2] if len(sys.argv) > 2 else "eager"
pose, losses = refine.refine_pose(init, target, H, W, FOCAL, kw, n_iters=5 if mode == "eager" else 12, graph=(mode != "eager"))
torch.cuda.synchronize()
print([round(float(l), 5) for l in losses])
