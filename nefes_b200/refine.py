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


class PoseRefiner:
    """The refinement iteration -- pose chain, render, loss, backward, Adam; ~250 launches, most of them few-element
    torch ops -- captured ONCE into a CUDA graph and replayed for every iteration of every query of the same shape.
    A query only rewrites the graph's static inputs (initial pose, target features) and zeroes the pose delta and the
    Adam state; the arithmetic is that of the eager loop (`refine_pose(..., graph=False)`)."""

    def __init__(self, H, W, focal, render_kwargs_test, lr_r, lr_t, chunk, device, feat_shape):
        self.args = (H, W, focal, chunk)
        self.kw = render_kwargs_test
        self.pose = LearnPose(1, True, True, torch.eye(4, device=device)[:3][None]).to(device)
        self.opt = torch.optim.Adam([{"params": [self.pose.r], "lr": lr_r}, {"params": [self.pose.t], "lr": lr_t}], capturable=True)
        self.target = torch.zeros(feat_shape, device=device)
        self.hist = torch.zeros(1, 10, device=device)
        self.graph, self.static_loss = None, None

    def _iter(self):
        H, W, focal, chunk = self.args
        c2w = self.pose(0)
        rgb, disp, acc, extras = render(H, W, focal, chunk=chunk, c2w=c2w[:3, :4], img_idx=self.hist, **self.kw)
        loss = feature_loss(extras["feat_map"].t(), self.target)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss.detach()

    @torch.no_grad()
    def _reset(self, init_c2w, feat_target, hist):
        self.pose.init_c2w[0].copy_(init_c2w[:3, :4])
        self.pose.r.zero_()
        self.pose.t.zero_()
        self.target.copy_(feat_target)
        if hist is not None:
            self.hist.copy_(hist)
        else:
            self.hist.zero_()
        for st in self.opt.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    def refine(self, init_c2w, feat_target, n_iters, hist=None):
        self._reset(init_c2w.to(self.target.device), feat_target, hist)
        losses, done = [], 0
        if self.graph is None:
            for _ in range(min(2, n_iters)):         # create the optimiser state, warm every kernel: real steps
                losses.append(self._iter())
                done += 1
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.static_loss = self._iter()
                self.graph = g
            except Exception:                        # capture is an optimisation; the eager loop is the definition
                self.graph = False
                torch.cuda.synchronize()
        for _ in range(n_iters - done):
            if self.graph:
                self.graph.replay()
                losses.append(self.static_loss.clone())
            else:
                losses.append(self._iter())
        with torch.no_grad():
            return self.pose(0)[:3, :4].clone(), losses


_REFINERS = {}


def refine_pose(init_c2w, feat_target, H, W, focal, render_kwargs_test, n_iters=50, lr_r=0.0087, lr_t=0.01,
                hist=None, chunk=32768, graph=None):
    """One query: `n_iters` Adam steps on the se(3)-style delta (DFM_pose_refine.py:380-440).
    feat_target [C, H*W].  Returns (refined c2w [3,4], list of losses).
    graph (default: on for n_iters >= 10 on CUDA): run the iterations as replays of a captured CUDA graph (PoseRefiner),
    cached per (camera, networks, learning rates) so that every further query pays no capture either."""
    dev = feat_target.device
    use_graph = ((n_iters >= 10) if graph is None else bool(graph)) and dev.type == "cuda"
    if use_graph:
        key = (H, W, float(focal), chunk, float(lr_r), float(lr_t), str(dev), tuple(feat_target.shape),
               id(render_kwargs_test.get("network_fn")), id(render_kwargs_test.get("network_fine")),
               getattr(render_kwargs_test.get("network_fn"), "precision", None),
               getattr(render_kwargs_test.get("network_fine"), "precision", None))
        ref = _REFINERS.get(key)
        if ref is None:
            ref = _REFINERS[key] = PoseRefiner(H, W, focal, render_kwargs_test, lr_r, lr_t, chunk, dev, tuple(feat_target.shape))
        return ref.refine(init_c2w, feat_target, n_iters, hist)
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
