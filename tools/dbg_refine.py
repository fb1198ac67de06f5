import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nefes_oracle as O
import nefes_b200 as nb
DEV="cuda"
H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.
w = np.load("tests/golden/weights.npz")
wc = {k[7:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("coarse/")}
wf = {k[5:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("fine/")}
c = nb.NeRFH_NFF("coarse", W=128); f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True)
c.load_state_dict(wc, strict=False); f.load_state_dict(wf); c.to(DEV); f.to(DEV)
for p in list(c.parameters()) + list(f.parameters()): p.requires_grad_(False)
class Args: nerfh_nff=True; use_fine_only=False; NeRFW=True; transient_at_test=True; netchunk=1<<21
q = lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR, perturb=0., raw_noise_std=0., test_time=True)
g = np.load("tests/golden/g5_render.npz")
c2w = torch.from_numpy(g["test/c2w"]).to(DEV).requires_grad_(True)
rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, c2w=c2w, img_idx=torch.zeros(1, 10), **kw)
sub = torch.from_numpy(g["test/sub"]).to(DEV)
loss = O.cosine_feature_loss(ex["feat_map"][sub].t(), torch.from_numpy(g["test/feat_target"]).to(DEV)) + rgb[sub].mean()
loss.backward()
print("gpu d_c2w\n", c2w.grad.cpu().numpy()); print("ref d_c2w\n", g["test/d_c2w"])
# refinement traces
h, w_, focal = 30, 40, FOCAL / 2
P = np.load("tests/golden/poses_stairs.npz")
gt = torch.tensor(P["test_gt"][0].reshape(3, 4), dtype=torch.float32); init = torch.tensor(P["dfnet_init"][0].reshape(3, 4), dtype=torch.float32)
with torch.no_grad():
    target = O.render(h, w_, focal, wc, wf, c2w=gt, near=NEAR, far=FAR, test_time=True)["feat_map"].t().contiguous()
from nefes_b200 import refine
r = torch.zeros(3, requires_grad=True); t = torch.zeros(3, requires_grad=True)
pose = refine.LearnPose(1, True, True, init[None].to(DEV)).to(DEV)
for it in range(4):
    c2 = O.learn_pose_c2w(r, t, init)
    out = O.render(h, w_, focal, wc, wf, c2w=c2, near=NEAR, far=FAR, test_time=True)
    l = O.cosine_feature_loss(out["feat_map"].t(), target); r.grad = t.grad = None; l.backward()
    cg = pose(0); rgb, disp, acc, ex = nb.render(h, w_, focal, chunk=32768, c2w=cg[:3, :4], img_idx=torch.zeros(1, 10), **kw)
    lg = refine.feature_loss(ex["feat_map"].t(), target.to(DEV)); pose.zero_grad(); lg.backward()
    print(it, "loss", float(l), float(lg), "\n  r.grad", r.grad.numpy(), pose.r.grad.cpu().numpy()[0], "\n  t.grad", t.grad.numpy(), pose.t.grad.cpu().numpy()[0])
    with torch.no_grad():
        step_r = -0.0087 * torch.sign(r.grad); step_t = -0.01 * torch.sign(t.grad)
        r += step_r; t += step_t; pose.r += step_r.to(DEV); pose.t += step_t.to(DEV)
