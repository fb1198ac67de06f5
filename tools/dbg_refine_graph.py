import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nefes_b200 as nb
from nefes_b200 import refine
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
c = nb.NeRFH_NFF("coarse", W=128).cuda(); f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True).cuda()
c.precision = f.precision = prec
for p in (c.flat, f.flat): p.requires_grad_(False)
H, W, focal = 60, 80, 65.688
g = np.load("tests/golden/poses_stairs.npz")
class Args: nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
q = lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=4., perturb=0., raw_noise_std=0., test_time=True)
gen = torch.Generator(device="cuda").manual_seed(11)
init = torch.tensor(g["dfnet_init"][0].reshape(3, 4), dtype=torch.float32, device="cuda")
target = torch.randn(128, H * W, device="cuda", generator=gen)
pe, le = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=False)
pe2, le2 = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=False)
pg, lg = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=True)
pg2, lg2 = refine.refine_pose(init, target, H, W, focal, kw, n_iters=12, graph=True)
print("eager :", [f"{float(x):.6f}" for x in le])
print("eager2:", [f"{float(x):.6f}" for x in le2])
print("graph :", [f"{float(x):.6f}" for x in lg])
print("graph2:", [f"{float(x):.6f}" for x in lg2])
print("pose eager-eager2", float((pe - pe2).abs().max()), "eager-graph", float((pe - pg).abs().max()), "graph-graph2", float((pg - pg2).abs().max()))
