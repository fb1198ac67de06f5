"""Test-time pose refinement through the renderer (`pose_only=3`):
script/test_refinement.py:74-96 -> dm/DFM_pose_refine.py:350-453 (loop), :290-348 (one step),
script/models/poses.py:25-50 (LearnPose), utils/lie_group_helper.py:60-81 (so(3) exp, the reference's
own lietorch=False path), dm/DFM_pose_refine.py:236-255 (cosine feature loss, per_pixel=False).

The render + its backward to the camera pose run on the engine; the 6-parameter pose chain, the loss
and Adam are the caller's few-element torch ops ("next" row).  FusionNet / exposure MLP /
DFNet is outside the path.  FusionNet and the affine colour transform (the stage right behind the render, 8f-2) run on the engine
too and join the iteration with `refine_pose(..., fusion=True)`; SURVEY.md 8d config C4 (the pinned benchmark definition) excludes them."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops, parallel
from .rendering import _fused_applies, render


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


def se3_exp(t, r):
    """SE3.exp([t, r]).matrix() of poses.py:31-32, 44 (the lietorch=True branch; lietorch itself is an un-vendored CUDA
    dependency of the reference, so this is the closed form it implements): R = Exp(r), translation = V(r) t with
    V = I + (1 - cos n)/n^2 K + (n - sin n)/n^3 K^2.  Evaluated in fp64 (series below n = 1e-4, where the closed-form
    coefficients cancel) and cast back; differentiable in t and r.  Returns (R [3,3], V t [3])."""
    dt = r.dtype
    r64, t64 = r.double(), t.double()
    K = vec2skew(r64)
    n2 = (r64 * r64).sum()
    n = torch.sqrt(n2.clamp_min(1e-300))
    small = n2 < 1e-8
    ns = torch.where(small, torch.ones_like(n), n)                    # keeps the unused branch finite for autograd
    a = torch.where(small, 1 - n2 / 6, torch.sin(ns) / ns)
    b = torch.where(small, 0.5 - n2 / 24, (1 - torch.cos(ns)) / ns ** 2)
    c = torch.where(small, 1 / 6 - n2 / 120, (ns - torch.sin(ns)) / ns ** 3)
    eye = torch.eye(3, dtype=torch.float64, device=r.device)
    KK = K @ K
    return (eye + a * K + b * KK).to(dt), ((eye + b * K + c * KK) @ t64).to(dt)


class LearnPose(nn.Module):
    """poses.py:6-50: c2w = [Exp(r) @ R0 | t + t0] per camera (lietorch=False, the reference's own pure-torch chain), or
    [Exp(r) @ R0 | V(r) t + t0] (lietorch=True: SE3.exp([t, r]), what dm/DFM_pose_refine.py:374 constructs)."""

    def __init__(self, num_cams, learn_R=True, learn_t=True, init_c2w=None, lietorch=False):
        super().__init__()
        self.num_cams = num_cams
        self.lietorch = bool(lietorch)
        self.init_c2w = None if init_c2w is None else nn.Parameter(init_c2w.clone(), requires_grad=False)
        self.r = nn.Parameter(torch.zeros(num_cams, 3), requires_grad=learn_R)
        self.t = nn.Parameter(torch.zeros(num_cams, 3), requires_grad=learn_t)

    def forward(self, cam_id: int):
        if self.lietorch:
            R, t = se3_exp(self.t[cam_id], self.r[cam_id])
        else:
            R, t = so3_exp(self.r[cam_id]), self.t[cam_id]
        if self.init_c2w is not None:
            R = R @ self.init_c2w[cam_id, :3, :3]
            t = t + self.init_c2w[cam_id, :3, 3]
        c2w = torch.eye(4, dtype=R.dtype, device=R.device)
        return torch.cat([torch.cat([R, t[:, None]], 1), c2w[3:]], 0)


def fix_coord_supp(args, pose, world_setup_dict, device=None):
    """dm/direct_pose_model.py:210-232: pose [N,3,4]; the translation column becomes ((x * pose_scale) + move_all_cam_vec) *
    pose_scale2.  Out of place (the reference writes into its argument; callers use the return value)."""
    if not torch.is_tensor(pose):
        pose = torch.as_tensor(pose, dtype=torch.float32, device=device)
    move = torch.as_tensor(world_setup_dict['move_all_cam_vec'], dtype=pose.dtype, device=pose.device)
    tr = (pose[:, :3, 3] * world_setup_dict['pose_scale'] + move) * world_setup_dict['pose_scale2']
    return torch.cat([pose[:, :3, :3], tr[..., None]], -1)


def svd_reg(pose):
    """dm/DFM_pose_refine.py:119-129 (Direct-PN: orthogonalise the rotation block): pose [B,3,4] -> [B,3,4], R <- U V^T."""
    u, s, v = torch.svd(pose[:, :3, :3])
    return torch.cat([u @ v.transpose(-2, -1), pose[:, :3, 3:]], -1)


def _chain6(lietorch, world_setup_dict):
    """The engine's pose-chain descriptor {se3, pose_scale, move(3), pose_scale2} (include/nefes_b200.h) as a ctypes array."""
    import ctypes
    w = world_setup_dict or {}
    mv = [float(x) for x in w.get('move_all_cam_vec', (0., 0., 0.))]
    return (ctypes.c_float * 6)(1.0 if lietorch else 0.0, float(w.get('pose_scale', 1.0)), mv[0], mv[1], mv[2],
                                float(w.get('pose_scale2', 1.0)))


def feature_loss(feature_rgb, feature_target):
    """DFM_pose_refine.py:236-255, img_in=False, per_pixel=False: inputs [C, N]."""
    return 1 - torch.nn.functional.cosine_similarity(feature_rgb, feature_target, dim=1, eps=1e-6).mean()


class _CosineLossPM(torch.autograd.Function):
    """1 - mean_c cos(feat[:, c], target[c, :]) on the engine (nefes_cosine_loss_{fwd,bwd}): feat [N,C] PIXEL-major (what the
    render / FusionNet / upsample kernels produce), target [C,N], optional pixel mask [N]."""

    @staticmethod
    def forward(ctx, feat, target, mask):
        L.need_cuda(feat, target, mask)
        f, t = L.f32c(feat), L.f32c(target)
        m = None if mask is None else L.f32c(mask.reshape(-1).float())
        N, C = f.shape
        if t.shape != (C, N) or C % 32 or C > 1024 or (m is not None and m.numel() != N):
            raise RuntimeError(f"nefes_b200: cosine loss expects feat [N,C], target [C,N] (C a multiple of 32, <= 1024), got {tuple(feat.shape)} / {tuple(target.shape)}")
        dev = f.device
        stats, loss, d_feat = torch.zeros(3, C, device=dev), torch.empty(1, device=dev), torch.empty_like(f)
        with torch.cuda.device(dev):
            st = L.stream_of(f)
            L.check(L.lib().nefes_cosine_loss_fwd(L.ptr(f), L.ptr(t), L.ptr(m), N, C, L.ptr(stats), st), "nefes_cosine_loss_fwd")
            L.check(L.lib().nefes_cosine_loss_bwd(L.ptr(f), L.ptr(t), L.ptr(m), L.ptr(stats), N, C, L.ptr(loss), None, None, 0, L.ptr(d_feat), st),
                    "nefes_cosine_loss_bwd")
        ctx.save_for_backward(d_feat)
        ctx.shape = tuple(feat.shape)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_feat,) = ctx.saved_tensors
        return (d_feat * g).reshape(ctx.shape), None, None


def cosine_loss_pm(feat_pm, target, mask=None):
    """feature_loss / masked_feature_loss (DFM_pose_refine.py:211-288, per_pixel=False) for a PIXEL-major rendered map:
    feat_pm [N,C], target [C,N] (or [C,H,W]), mask [N] / [1,H,W] or None.  Engine kernels; differentiable in feat_pm."""
    return _CosineLossPM.apply(feat_pm, target.reshape(target.shape[0], -1), mask)


def masked_feature_loss(feature_rgb, feature_target, mask, img_in=True, per_pixel=False):
    """Drop-in for DFM_pose_refine.py:257-288 (channel-major inputs [C,H,W] or [C,N], as the reference passes them)."""
    if per_pixel:
        raise RuntimeError("nefes_b200: per_pixel=True is not built (the reference calls per_pixel=False)")
    C = feature_rgb.shape[0]
    fr = feature_rgb.reshape(C, -1)
    if fr.is_cuda:
        return cosine_loss_pm(fr.t().contiguous(), feature_target.reshape(C, -1), mask)
    valid = torch.nonzero(mask.reshape(-1) > 0, as_tuple=True)[0]
    return 1 - torch.nn.functional.cosine_similarity(fr[:, valid], feature_target.reshape(C, -1)[:, valid], dim=1, eps=1e-6).mean()


class _UpsampleCrop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, w, H, W, crop):
        L.need_cuda(x)
        xc = L.f32c(x)
        C = xc.shape[1]
        if xc.shape[0] != h * w:
            raise RuntimeError(f"nefes_b200: upsample_crop expects [h*w, C] = [{h * w}, C], got {tuple(x.shape)}")
        out = torch.empty((H - 2 * crop) * (W - 2 * crop), C, device=xc.device)
        with torch.cuda.device(xc.device):
            L.check(L.lib().nefes_upsample_crop_fwd(L.ptr(xc), h, w, C, H, W, crop, L.ptr(out), L.stream_of(xc)), "nefes_upsample_crop_fwd")
        ctx.meta = (h, w, C, H, W, crop, tuple(x.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        h, w, C, H, W, crop, shape = ctx.meta
        g = L.f32c(g)
        tmp, dx = torch.empty(h * (W - 2 * crop) * C, device=g.device), torch.empty(h * w, C, device=g.device)
        with torch.cuda.device(g.device):
            L.check(L.lib().nefes_upsample_crop_bwd(L.ptr(g), h, w, C, H, W, crop, L.ptr(tmp), L.ptr(dx), L.stream_of(g)), "nefes_upsample_crop_bwd")
        return dx.reshape(shape), None, None, None, None, None


def upsample_crop(x_pm, h, w, H, W, crop=10):
    """torch.nn.Upsample(size=(H, W), mode='bicubic') + [:, :, crop:-crop, crop:-crop] (DFM_APR_refine.py:114-124) of a
    pixel-major map x_pm [h*w, C] -> [(H - 2 crop) * (W - 2 crop), C]; engine kernels, differentiable."""
    return _UpsampleCrop.apply(x_pm, int(h), int(w), int(H), int(W), int(crop))


def apr_feature_loss(feature_pm, feature_target, h, w, crop=10, mask=None):
    """The loss of the APR-refinement step (DFM_APR_refine.py:114-129): the fused feature map rendered at h x w is up-sampled
    bicubically to the target's H x W, both are cropped by `crop` pixels, cosine feature loss (optionally masked).
    feature_pm [h*w, C] pixel-major (FusionNet output), feature_target [C,H,W]."""
    C, H, W = feature_target.shape
    up = upsample_crop(feature_pm, h, w, H, W, crop)
    tgt = feature_target[:, crop:H - crop, crop:W - crop].reshape(C, -1)
    m = None if mask is None else mask.reshape(H, W)[crop:H - crop, crop:W - crop].reshape(-1)
    return cosine_loss_pm(up, tgt, m)


class _EncodeHist:
    encode_hist = True


def _fused_features(kw, rgb, feat_map, hist, H, W, fusion, encode_hist):
    """What the loss sees ([C, H*W]).  fusion=False: the rendered feature map (the pinned definition).
    fusion=True: the reference's full step, DFM_pose_refine.py:324-329 -- rgb -> affine_color_transform (args.encode_hist) ->
    FusionNet(rgb, feat_map) -- so the pose gradient also flows through the rendered colours."""
    if not fusion:
        return feat_map.t()
    net = kw["network_fn"]
    if encode_hist:
        rgb = net.affine_color_transform(_EncodeHist(), rgb, hist, 1)
    return net.run_fusion_net(rgb, feat_map, H, W, 1)[2][0].reshape(feat_map.shape[1], -1)


class PoseRefiner:
    """The refinement iteration -- pose chain, render, loss, backward, Adam; ~250 launches, most of them few-element
    torch ops -- captured ONCE into a CUDA graph and replayed for every iteration of every query of the same shape.
    A query only rewrites the graph's static inputs (initial pose, target features) and zeroes the pose delta and the
    Adam state; the arithmetic is that of the eager loop (`refine_pose(..., graph=False)`)."""

    def __init__(self, H, W, focal, render_kwargs_test, lr_r, lr_t, chunk, device, feat_shape, lietorch=False, world_setup_dict=None,
                 fusion=False, encode_hist=False):
        self.args = (H, W, focal, chunk)
        self.kw = render_kwargs_test
        self.world = world_setup_dict
        self.fusion, self.encode_hist = bool(fusion), bool(encode_hist)
        self.pose = LearnPose(1, True, True, torch.eye(4, device=device)[:3][None], lietorch=lietorch).to(device)
        self.opt = torch.optim.Adam([{"params": [self.pose.r], "lr": lr_r}, {"params": [self.pose.t], "lr": lr_t}], capturable=True)
        self.target = torch.zeros(feat_shape, device=device)
        self.hist = torch.zeros(1, 10, device=device)
        self.graph, self.static_loss = None, None

    def _iter(self):
        H, W, focal, chunk = self.args
        c2w = self.pose(0)[:3, :4]
        if self.world is not None:
            c2w = fix_coord_supp(None, c2w[None], self.world)[0]
        rgb, disp, acc, extras = render(H, W, focal, chunk=chunk, c2w=c2w, img_idx=self.hist, **self.kw)
        loss = feature_loss(_fused_features(self.kw, rgb, extras["feat_map"], self.hist, H, W, self.fusion, self.encode_hist), self.target)
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
            except Exception as e:                   # capture is an optimisation; the eager loop is the definition
                import warnings
                warnings.warn(f"nefes_b200: CUDA-graph capture of the refinement iteration failed ({e}); running eagerly")
                self.graph = False
                torch.cuda.synchronize()
        for _ in range(n_iters - done):
            if self.graph:
                self.graph.replay()
                losses.append(self.static_loss.clone())
            else:
                losses.append(self._iter())
        with torch.no_grad():
            return self.pose(0)[:3, :4].clone(), losses        # the LEARNED pose (before fix_coord_supp), as the reference reports it


class EnginePoseRefiner:
    """The whole refinement iteration as engine launches (24 kernels, no torch op): pose chain + camera rays + packing
    (nefes_pose_rays_fwd), render_rays forward (nefes_render_rays_fwd), cosine feature loss and its gradient
    (nefes_cosine_loss_*), render_rays backward to the rays, rays -> pose cotangent (nefes_pose_rays_bwd), so(3) chain +
    Adam (nefes_pose_adam_step).  Captured once into a CUDA graph and replayed; the loss of iteration k lands in
    `loss_hist[k]` on the device, so a query is n_iters graph launches and nothing else.  Applies to the reference's own
    refinement configuration (standard query function, frozen fields, test_time render, per_pixel=False loss)."""
    HIST = 1024

    def __init__(self, H, W, focal, kw, lr_r, lr_t, device, n_ch, lietorch=False, world_setup_dict=None):
        c, f, args = kw["network_fn"], kw["network_fine"], kw["args"]
        self.chain, self.plain = _chain6(lietorch, world_setup_dict), _chain6(lietorch, None)
        dev = torch.device(device)
        self.H, self.W, self.focal, self.near, self.far = int(H), int(W), float(focal), float(kw["near"]), float(kw["far"])
        self.lr_r, self.lr_t, self.N, self.C, self.dev = float(lr_r), float(lr_t), int(H) * int(W), int(n_ch), dev
        from .nerfh_nff import _PREC
        cfg = dict(n_samples=int(kw["N_samples"]), n_importance=int(kw["N_importance"]), prec=_PREC[c.precision], test_time=True,
                   output_transient=bool(args.NeRFW), transient_at_test=bool(args.transient_at_test), net_coarse=c.net_id,
                   net_fine=f.net_id, beta_min=f.beta_min)
        self.call = ops.RenderCall(self.N, 21, cfg, c.flat, f.flat, dev, frozen_weights=True)   # packed once per query (refine())
        z = lambda *shape: torch.zeros(*shape, device=dev)
        self.pose6, self.init, self.c2w, self.d_c2w = z(6), z(3, 4), z(3, 4), z(12)
        self.state, self.stats = z(13), z(3, self.C)
        self.target, self.g_feat = z(self.C, self.N), z(self.N, self.C)
        self.loss, self.loss_hist = z(1), z(self.HIST)
        self.graph = None

    def _iter(self):
        lib, p, st = L.lib(), L.ptr, L.stream_of(self.pose6)
        with torch.cuda.device(self.dev):
            L.check(lib.nefes_pose_rays_fwd(p(self.pose6), p(self.init), self.H, self.W, self.focal, self.near, self.far,
                                            p(self.c2w), p(self.call.rays), 21, self.chain, st), "nefes_pose_rays_fwd")
            self.call.forward()
            L.check(lib.nefes_cosine_loss_fwd(p(self.call.feat), p(self.target), None, self.N, self.C, p(self.stats), st),
                    "nefes_cosine_loss_fwd")
            L.check(lib.nefes_cosine_loss_bwd(p(self.call.feat), p(self.target), None, p(self.stats), self.N, self.C, p(self.loss),
                                              p(self.loss_hist), p(self.state[12:]), self.HIST, p(self.g_feat), st),
                    "nefes_cosine_loss_bwd")
            d_rays = self.call.backward(g_feat=self.g_feat)
            L.check(lib.nefes_pose_rays_bwd(p(d_rays), p(self.call.rays), 21, self.H, self.W, self.focal, p(self.d_c2w), st),
                    "nefes_pose_rays_bwd")
            L.check(lib.nefes_pose_adam_step(p(self.pose6), p(self.init), p(self.d_c2w), p(self.stats), 3 * self.C,
                                             p(self.state), self.lr_r, self.lr_t, 0.9, 0.999, 1e-8, self.chain, st), "nefes_pose_adam_step")

    @torch.no_grad()
    def refine(self, init_c2w, feat_target, n_iters, use_graph=True):
        if n_iters > self.HIST:
            raise RuntimeError(f"nefes_b200: at most {self.HIST} refinement iterations per query")
        self.init.copy_(init_c2w[:3, :4])
        self.target.copy_(feat_target)
        for t in (self.pose6, self.state, self.d_c2w, self.stats):
            t.zero_()
        self.call.prepack()                           # the fields are frozen for the query: their operand images are built here, once
        done = 0
        if use_graph and self.graph is None:
            for _ in range(min(2, n_iters)):         # warm every kernel with real steps, then capture one iteration
                self._iter()
                done += 1
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._iter()
            self.graph = g
        for _ in range(n_iters - done):
            if use_graph:
                self.graph.replay()
            else:
                self._iter()
        L.check(L.lib().nefes_pose_rays_fwd(L.ptr(self.pose6), L.ptr(self.init), self.H, self.W, self.focal, self.near, self.far,
                                            L.ptr(self.c2w), L.ptr(self.call.rays), 21, self.plain, L.stream_of(self.pose6)),
                "nefes_pose_rays_fwd")               # learned c2w of the final parameters (before fix_coord_supp)
        return self.c2w.clone(), list(self.loss_hist[:n_iters].clone())


def _engine_refiner_applies(feat_target, H, W, kw, hist):
    c, f, args = kw.get("network_fn"), kw.get("network_fine"), kw.get("args")
    if feat_target.dim() != 2 or feat_target.shape[1] != H * W or feat_target.shape[0] % 32 or feat_target.shape[0] > 1024:
        return False
    if not kw.get("test_time") or kw.get("perturb") or kw.get("raw_noise_std") or kw.get("white_bkgd") or kw.get("lindisp"):
        return False
    if not kw.get("use_viewdirs") or kw.get("ndc", True) or not getattr(args, "nerfh_nff", False):
        return False
    if c is None or f is None or c.flat.requires_grad or f.flat.requires_grad:
        return False
    q = kw.get("network_query_fn")
    if H * W * (int(kw.get("N_samples") or 0) + int(kw.get("N_importance") or 0)) > getattr(q, "netchunk", 0):
        return False                                   # the iteration's RenderCall is ONE engine call
    probe = torch.empty(H * W, 21, device="meta")
    return _fused_applies(probe, c, kw.get("network_query_fn"), kw.get("N_samples"), kw.get("N_importance", 0), f, args)


_REFINERS = {}
_MAX_REFINERS = 8


def clear_refiner_cache():
    """Drop every cached refiner (captured graphs + workspaces)."""
    _REFINERS.clear()


def refine_pose(init_c2w, feat_target, H, W, focal, render_kwargs_test, n_iters=50, lr_r=0.0087, lr_t=0.01,
                hist=None, chunk=32768, graph=None, engine=None, lietorch=False, world_setup_dict=None, fusion=False,
                encode_hist=False):
    """One query: `n_iters` Adam steps on the se(3)-style delta (DFM_pose_refine.py:380-440).
    feat_target [C, H*W].  Returns (refined c2w [3,4], list of losses).
    graph (default: on for n_iters >= 10 on CUDA): run the iterations as replays of a captured CUDA graph (PoseRefiner),
    cached per (camera, networks, learning rates) so that every further query pays no capture either.
    engine (default: on where it applies): the iteration is EnginePoseRefiner's 24 engine launches (no torch op); False
    keeps the torch pose chain / loss / optimiser around the engine render; True raises where it does not apply.
    lietorch: the pose delta goes through SE3.exp([t, r]) (poses.py:31-32, what DFM_pose_refine.py:374 constructs) instead of
    the reference's pure-torch [Exp(r) | t].  world_setup_dict: fix_coord_supp's pose_scale / move_all_cam_vec / pose_scale2
    (direct_pose_model.py:210-232) between the learned pose and the renderer; the returned pose is the learned one.
    fusion: the loss is taken on FusionNet(affine_color_transform(rgb), feat_map) as DFM_pose_refine.py:324-337 does on nerfh_nff
    configs (feat_target is then the DFNet feature map); the engine-only iteration does not cover it, the graph-replayed
    torch-glue iteration (render, FusionNet and colour transform on the engine) does."""
    dev = feat_target.device
    use_graph = ((n_iters >= 10) if graph is None else bool(graph)) and dev.type == "cuda"
    kw_ = render_kwargs_test
    nets = (kw_.get("network_fn"), kw_.get("network_fine"))
    a_ = kw_.get("args")
    # everything a cached refiner bakes in: camera, rates, sample counts, flags, and the STORAGE of the frozen weights
    # (a refiner holds raw pointers to flat.detach(): model.to() / a reloaded Parameter must miss the cache)
    key = (H, W, float(focal), chunk, float(lr_r), float(lr_t), str(dev), tuple(feat_target.shape),
           tuple(id(n) for n in nets), tuple(getattr(n, "precision", None) for n in nets),
           tuple(n.flat.data_ptr() if hasattr(n, "flat") else None for n in nets),
           kw_.get("N_samples"), kw_.get("N_importance"), getattr(kw_.get("network_query_fn"), "netchunk", None),
           bool(getattr(a_, "NeRFW", False)), bool(getattr(a_, "transient_at_test", False)),
           bool(getattr(a_, "nerfh_nff", False)), bool(getattr(a_, "use_fine_only", False)),
           float(kw_.get("near", 0.)), float(kw_.get("far", 1.)), bool(lietorch), tuple(_chain6(lietorch, world_setup_dict)),
           bool(fusion), bool(encode_hist))
    while len(_REFINERS) >= _MAX_REFINERS:           # bounded: a refiner pins its GPU workspaces
        _REFINERS.pop(next(iter(_REFINERS)))
    if (engine is None or engine) and not fusion and dev.type == "cuda" and H * W <= chunk and \
            _engine_refiner_applies(feat_target, H, W, render_kwargs_test, hist):
        ref = _REFINERS.get(("engine",) + key)
        if ref is None:
            ref = _REFINERS[("engine",) + key] = EnginePoseRefiner(H, W, focal, render_kwargs_test, lr_r, lr_t, dev,
                                                                   feat_target.shape[0], lietorch, world_setup_dict)
        return ref.refine(init_c2w.to(dev), feat_target, n_iters, use_graph=use_graph)
    if engine:
        raise RuntimeError("nefes_b200: the engine-resident refinement iteration does not cover this configuration "
                           "(needs StandardQuery, frozen fields, test_time render, feat_target [C, H*W])")
    if use_graph:
        ref = _REFINERS.get(key)
        if ref is None:
            ref = _REFINERS[key] = PoseRefiner(H, W, focal, render_kwargs_test, lr_r, lr_t, chunk, dev, tuple(feat_target.shape),
                                               lietorch, world_setup_dict, fusion, encode_hist)
        return ref.refine(init_c2w, feat_target, n_iters, hist)
    pose = LearnPose(1, True, True, init_c2w[None].to(dev), lietorch=lietorch).to(dev)
    opt = torch.optim.Adam([{"params": [pose.r], "lr": lr_r}, {"params": [pose.t], "lr": lr_t}])
    hist = torch.zeros(1, 10, device=dev) if hist is None else hist
    losses = []
    for _ in range(n_iters):
        c2w = pose(0)[:3, :4]
        if world_setup_dict is not None:
            c2w = fix_coord_supp(None, c2w[None], world_setup_dict)[0]
        rgb, disp, acc, extras = render(H, W, focal, chunk=chunk, c2w=c2w, img_idx=hist, **render_kwargs_test)
        loss = feature_loss(_fused_features(render_kwargs_test, rgb, extras["feat_map"], hist, H, W, fusion, encode_hist), feat_target)
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
