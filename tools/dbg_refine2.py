import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nefes_oracle as O
import nefes_b200 as nb
from nefes_b200 import refine
DEV="cuda"
FOCAL, NEAR, FAR = 525.505 / 2 / 4, 0., 4.
w = np.load("tests/golden/weights.npz")
wc = {k[7:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("coarse/")}
wf = {k[5:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("fine/")}
c = nb.NeRFH_NFF("coarse", W=128); f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True)
c.load_state_dict(wc, strict=False); f.load_state_dict(wf); c.to(DEV); f.to(DEV)
for p in list(c.parameters()) + list(f.parameters()): p.requires_grad_(False)
class Args: nerfh_nff=True; use_fine_only=False; NeRFW=True; transient_at_test=True; netchunk=1<<21
q = lambda i, v, ts, fn, typ, ot, test_time, store_rgb: nb.run_network_NeRFH_NFF(i, v, ts, fn, typ=typ, output_transient=ot, netchunk=Args.netchunk, test_time=test_time, store_rgb=store_rgb)
kw = dict(network_query_fn=q, N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR, perturb=0., raw_noise_std=0., test_time=True)
h, w_, focal = 30, 40, FOCAL / 2
P = np.load("tests/golden/poses_stairs.npz")
gt = torch.tensor(P["test_gt"][0].reshape(3, 4), dtype=torch.float32); init = torch.tensor(P["dfnet_init"][0].reshape(3, 4), dtype=torch.float32)
with torch.no_grad():
    target = O.render(h, w_, focal, wc, wf, c2w=gt, near=NEAR, far=FAR, test_time=True)["feat_map"].t().contiguous()
Pc64, Pf64 = O.clone_params(wc, torch.float64), O.clone_params(wf, torch.float64)
r64 = torch.zeros(3, dtype=torch.float64, requires_grad=True); t64 = torch.zeros(3, dtype=torch.float64, requires_grad=True)
opt = torch.optim.Adam([{"params": [r64], "lr": 0.0087}, {"params": [t64], "lr": 0.01}])
for it in range(8):
    c2 = O.learn_pose_c2w(r64, t64, init.double())
    out = O.render(h, w_, focal, Pc64, Pf64, c2w=c2, near=NEAR, far=FAR, test_time=True, hist=torch.zeros(1, 10, dtype=torch.float64), return_aux=True)
    l = O.cosine_feature_loss(out["feat_map"].t(), target.double()); opt.zero_grad(); l.backward()
    # fp32 reference at the same pose
    r32 = r64.detach().float().requires_grad_(True); t32 = t64.detach().float().requires_grad_(True)
    o32 = O.render(h, w_, focal, wc, wf, c2w=O.learn_pose_c2w(r32, t32, init), near=NEAR, far=FAR, test_time=True, return_aux=True)
    O.cosine_feature_loss(o32["feat_map"].t(), target).backward()
    pose = refine.LearnPose(1, True, True, init[None].to(DEV)).to(DEV)
    with torch.no_grad(): pose.r.copy_(r64.detach().float()[None]); pose.t.copy_(t64.detach().float()[None])
    rgb, disp, acc, ex = nb.render(h, w_, focal, chunk=32768, c2w=pose(0)[:3, :4], img_idx=torch.zeros(1, 10), return_aux=True, **kw)
    refine.feature_loss(ex["feat_map"].t(), target.to(DEV)).backward()
    e = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
    inds_mis = float((ex["aux_inds"].cpu() != out["_aux"]["inds"].int()).float().mean()); inds_mis32 = float((o32["_aux"]["inds"] != out["_aux"]["inds"]).float().mean())
    print(f"it {it}: t.grad64 {t64.grad.numpy()} | rel err t: engine {e(pose.t.grad[0], t64.grad):.2e} ref32 {e(t32.grad, t64.grad):.2e} | r: engine {e(pose.r.grad[0], r64.grad):.2e} ref32 {e(r32.grad, r64.grad):.2e} | feat err engine {e(ex['feat_map'], out['feat_map']):.2e} ref32 {e(o32['feat_map'], out['feat_map']):.2e} | inds mismatch engine {inds_mis:.4f} ref32 {inds_mis32:.4f}")
    opt.step()
