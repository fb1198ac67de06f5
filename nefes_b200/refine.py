"""Test-time pose refinement through the renderer (`pose_only=3`):
script/test_refinement.py:74-96 -> dm/DFM_pose_refine.py:350-453 (loop), :290-348 (one step),
script/models/poses.py:25-50 (LearnPose), utils/lie_group_helper.py:60-81 (so(3) exp, the reference's
own lietorch=False path), dm/DFM_pose_refine.py:236-255 (cosine feature loss, per_pixel=False).

The render + its backward to the camera pose run on the engine; the 6-parameter pose chain, the loss
and Adam are the caller's few-element torch ops ("next" row).  FusionNet / exposure MLP /
DFNet are outside the path and excluded, as stated in SURVEY.md 8d config C4."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import parallel
from .rendering import render


def vec2skew(v):
    zero = torch.zeros(1, dtype=v.dtype, device=v.device)
    return torch.stack([torch.cat([zero, -v[2:3], v[1:2]]), torch.cat([v[2:3], zero, -v[0:1]]),
                        torch.cat([-v[1:2], v[0:1], zero])], dim=0)


def so3_exp(r):
    """lie_group_helper.py:60-70."""
    K = vec2skew(r)
    n = r.norm() + 1e-15
    eye = torch.eye(3, dtype=r.dtype, device=r.device)
    return eye + (torch.sin(n) / n) * K + ((1 - torch.cos(n)) / n ** 2) * (K @ K)


class LearnPose(nn.Module):
    """poses.py:6-50 with lietorch=False: c2w = [Exp(r) @ R0 | t + t0] per camera."""

    def __init__(self, num_cams, learn_R=True, learn_t=True, init_c2w=None, lietorch=False):
        super().__init__()
        if lietorch:
            raise RuntimeError("nefes_b200: the lietorch SE3 path is a third-party CUDA dependency that is not vendored; "
                               "use lietorch=False (the reference's own pure-torch path)")
        self.num_cams = num_cams
        self.init_c2w = None if init_c2w is None else nn.Parameter(init_c2w.clone(), requires_grad=False)
        self.r = nn.Parameter(torch.zeros(num_cams, 3), requires_grad=learn_R)
        self.t = nn.Parameter(torch.zeros(num_cams, 3), requires_grad=learn_t)

    def forward(self, cam_id: int):
        R = so3_exp(self.r[cam_id])
        t = self.t[cam_id]
        if self.init_c2w is not None:
            R = R @ self.init_c2w[cam_id, :3, :3]
            t = t + self.init_c2w[cam_id, :3, 3]
        c2w = torch.eye(4, dtype=R.dtype, device=R.device)
        return torch.cat([torch.cat([R, t[:, None]], 1), c2w[3:]], 0)


def feature_loss(feature_rgb, feature_target):
    """DFM_pose_refine.py:236-255, img_in=False, per_pixel=False: inputs [C, N]."""
    return 1 - torch.nn.functional.cosine_similarity(feature_rgb, feature_target, dim=1, eps=1e-6).mean()


def refine_pose(init_c2w, feat_target, H, W, focal, render_kwargs_test, n_iters=50, lr_r=0.0087, lr_t=0.01,
                hist=None, chunk=32768):
    """One query: `n_iters` Adam steps on the se(3)-style delta (DFM_pose_refine.py:380-440).
    feat_target [C, H*W].  Returns (refined c2w [3,4], list of losses)."""
    dev = feat_target.device
    pose = LearnPose(1, True, True, init_c2w[None].to(dev)).to(dev)
    opt = torch.optim.Adam([{"params": [pose.r], "lr": lr_r}, {"params": [pose.t], "lr": lr_t}])
    hist = torch.zeros(1, 10, device=dev) if hist is None else hist
    losses = []
    for _ in range(n_iters):
        c2w = pose(0)
        rgb, disp, acc, extras = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], img_idx=hist, **render_kwargs_test)
        loss = feature_loss(extras["feat_map"].t(), feat_target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    with torch.no_grad():
        return pose(0)[:3, :4].clone(), losses


def refine_queries(init_c2ws, feat_targets, H, W, focal, render_kwargs_test, **kw):
    """Queries sharded over ranks (r, r+W, ...), no collective until the final gather of [n,12] poses
    (SURVEY.md 8e)."""
    rank, ws = parallel.world()
    n = init_c2ws.shape[0]
    ids = parallel.shard_strided(n, rank, ws)
    out = [refine_pose(init_c2ws[i], feat_targets[i], H, W, focal, render_kwargs_test, **kw)[0].reshape(12) for i in ids]
    local = torch.stack(out) if out else torch.zeros(0, 12, device=init_c2ws.device)
    return parallel.gather_rows(local, ids, n).reshape(n, 3, 4)
