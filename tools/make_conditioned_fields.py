"""GPU: train the two fields for a couple of thousand engine steps on a synthetic, pose-dependent target so that rendered
colours and features CARRY SIGNAL about the camera pose (random-init fields render an almost constant image, and a pose
gradient at fp32-noise level makes any refined-pose comparison meaningless -- VERDICT r1, weak 1).  Writes
tests/golden/c4_fields.npz (state_dict tensors of both fields, fp32): the conditioned C4 problem of
oracle/make_c4_fixture.py and tests/test_gpu_refine_c4.py.  Run once on a GPU box:  python tools/make_conditioned_fields.py

Scene: a ray (o, d) sees the point x = o + 2 d/|d|; rgb(x) = 0.5 + 0.5 sin(1.5 x + phase), feature_c(x) = sin(k_c . x + phi_c)
with k_c ~ N(0, 1.2^2).  Poses: the first 7-Scenes-stairs test pose perturbed by up to 0.45 m / 8 degrees."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import nefes_b200 as nb
from nefes_b200 import refine

H, W, FOCAL, NEAR, FAR = 60, 80, 525.505 / 2 / 4, 0., 4.
dev = torch.device("cuda")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "poses_stairs.npz"))
gt = torch.tensor(g["test_gt"][0].reshape(3, 4), dtype=torch.float32, device=dev)
gen = torch.Generator(device=dev).manual_seed(2026)
kf = torch.randn(3, 128, device=dev, generator=gen) * 1.2
pf = torch.rand(128, device=dev, generator=gen) * 6.2832
pc = torch.tensor([0.0, 2.1, 4.2], device=dev)


def scene(o, d):
    x = o + 2.0 * d / d.norm(dim=-1, keepdim=True)
    return 0.5 + 0.5 * torch.sin(1.5 * x + pc), torch.sin(x @ kf + pf)


class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
coarse = nb.NeRFH_NFF("coarse", W=128, precision="bf16").to(dev)
fine = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True, precision="bf16").to(dev)
kw = dict(network_query_fn=nb.StandardQuery(Args.netchunk), N_importance=64, N_samples=64, network_fn=coarse, network_fine=fine,
          use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=NEAR, far=FAR, perturb=1., raw_noise_std=0.,
          test_time=False, retraw=True)
opt = nb.FlatAdam([coarse.flat, fine.flat], lr=1e-3)
loss_fn = nb.ColorFeatureFusionNerfWLoss(coef=1, L1_loss=True)
B, NR = 4, 1536
for it in range(steps):
    r = (torch.rand(B, 3, device=dev, generator=gen) - 0.5) * 2 * np.radians(8.0)
    t = (torch.rand(B, 3, device=dev, generator=gen) - 0.5) * 2 * 0.45
    poses = torch.stack([torch.cat([refine.so3_exp(r[b]) @ gt[:, :3], (gt[:, 3] + t[b])[:, None]], 1) for b in range(B)])
    ro, rd = nb.get_rays_batch(H, W, FOCAL, poses)
    idx = torch.stack([torch.randperm(H * W, device=dev, generator=gen)[:NR] for _ in range(B)])
    ro = torch.gather(ro.reshape(B, -1, 3), 1, idx[..., None].expand(-1, -1, 3)).reshape(-1, 3)
    rd = torch.gather(rd.reshape(B, -1, 3), 1, idx[..., None].expand(-1, -1, 3)).reshape(-1, 3)
    tgt_rgb, tgt_f = scene(ro, rd)
    rgb, disp, acc, ex = nb.render(H, W, FOCAL, chunk=32768, rays=(ro, rd), img_idx=torch.zeros(B * NR, 10, device=dev), **kw)
    res = {"rgb_coarse": ex["rgb0"], "rgb_fine": rgb, "beta": ex["beta"], "transient_sigmas": ex["transient_sigmas"], "feat_fine": ex["feat_map"],
           "feat_coarse": ex["feat0"]}
    l_rgb, l_f = loss_fn(res, {"rgb": tgt_rgb, "feat": tgt_f}, switch_on=False, color_only_switch=False)
    loss = l_rgb + 0.5 * l_f
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    if it % 250 == 0 or it == steps - 1:
        print(f"step {it}: loss {float(loss):.4f}  colour {float(l_rgb):.4f}  feature L1 {float(l_f):.4f}  feat std over rays {float(ex['feat_map'].std(0).mean()):.4f}", flush=True)
out = {}
for name, net in (("coarse", coarse), ("fine", fine)):
    for k, v in net.state_dict().items():
        if k.startswith("fusion_net") or k.startswith("exposure"):
            continue
        out[f"{name}/{k}"] = v.detach().float().cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/c4_fields.npz", **out)
print("wrote gpurun_out/c4_fields.npz", sum(v.size for v in out.values()), "floats")
