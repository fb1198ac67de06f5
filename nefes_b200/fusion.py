"""Post-render appearance + fusion stage on the engine: autograd wrappers over nefes_fusion_{fwd,bwd} and
nefes_affine_color_{fwd,bwd} (csrc/fusion.cu).  The torch modules (FusionNet, ExposureMLP in nerfh_nff.py) remain the
owners of the parameters under the reference's state_dict keys; on CUDA their forward goes through these functions."""
import ctypes

import torch
from torch.autograd import Function

from . import _lib as L


class _FusionParams(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_void_p * 4), ("bias", ctypes.c_void_p * 4), ("bn_weight", ctypes.c_void_p),
                ("bn_bias", ctypes.c_void_p), ("bn_running_mean", ctypes.c_void_p), ("bn_running_var", ctypes.c_void_p)]


class _FusionGrads(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_void_p * 4), ("bias", ctypes.c_void_p * 4), ("bn_weight", ctypes.c_void_p), ("bn_bias", ctypes.c_void_p)]


def _params_struct(ws, bs, bn):
    p = _FusionParams()
    for i in range(4):
        p.weight[i], p.bias[i] = ws[i].data_ptr(), bs[i].data_ptr()
    if bn is not None:
        p.bn_weight, p.bn_bias, p.bn_running_mean, p.bn_running_var = (t.data_ptr() for t in bn)
    return p


class _FusionNetFn(Function):
    """out [P,128] = FusionNet(rgb [P,3], feat [P,128]); parameters: 4 conv weights, 4 biases, (BN weight, bias)."""

    @staticmethod
    def forward(ctx, rgb, feat, B, H, W, training, residual, momentum, eps, run_mean, run_var, tf32, *params):
        L.need_cuda(rgb, feat, *params)
        no_bn = len(params) == 8
        rgb_c, feat_c = L.f32c(rgb), L.f32c(feat)
        ps = [L.f32c(p) for p in params]
        P = B * H * W
        if rgb_c.shape != (P, 3) or feat_c.shape != (P, 128):
            raise RuntimeError(f"nefes_b200: fusion net expects rgb [{P},3] and features [{P},128], got {tuple(rgb.shape)} / {tuple(feat.shape)}")
        dev = rgb_c.device
        ws = torch.empty(int(L.lib().nefes_fusion_workspace(P)) // 4, device=dev)
        out = torch.empty(P, 128, device=dev)
        bn = None if no_bn else (ps[8], ps[9], run_mean, run_var)
        st = _params_struct(ps[0:4], ps[4:8], bn)
        with torch.cuda.device(dev), _GemmMode(tf32):
            L.check(L.lib().nefes_fusion_fwd(ctypes.byref(st), L.ptr(rgb_c), L.ptr(feat_c), B, H, W, int(training), int(no_bn), int(residual),
                                             float(momentum), float(eps), L.ptr(ws), L.ptr(out), L.stream_of(rgb_c)), "nefes_fusion_fwd")
        ctx.save_for_backward(ws, *ps)
        ctx.meta = (B, H, W, int(training), int(no_bn), int(residual), run_mean, run_var, tuple(rgb.shape), tuple(feat.shape), tf32)
        return out

    @staticmethod
    def backward(ctx, g):
        ws, *ps = ctx.saved_tensors
        B, H, W, training, no_bn, residual, run_mean, run_var, s_rgb, s_feat, tf32 = ctx.meta
        P = B * H * W
        dev = ws.device
        g = L.f32c(g)
        need = ctx.needs_input_grad
        d_rgb = torch.empty(P, 3, device=dev) if need[0] else None
        d_feat = torch.empty(P, 128, device=dev) if need[1] else None
        grads = [torch.zeros_like(p) if need[12 + i] else None for i, p in enumerate(ps)]
        gs = _FusionGrads()
        for i in range(4):
            gs.weight[i] = grads[i].data_ptr() if grads[i] is not None else None
            gs.bias[i] = grads[4 + i].data_ptr() if grads[4 + i] is not None else None
        if not no_bn:
            gs.bn_weight = grads[8].data_ptr() if grads[8] is not None else None
            gs.bn_bias = grads[9].data_ptr() if grads[9] is not None else None
        bn = None if no_bn else (ps[8], ps[9], run_mean, run_var)
        st = _params_struct(ps[0:4], ps[4:8], bn)
        with torch.cuda.device(dev), _GemmMode(tf32):
            L.check(L.lib().nefes_fusion_bwd(ctypes.byref(st), ctypes.byref(gs), L.ptr(g), B, H, W, training, no_bn, residual, L.ptr(ws),
                                             L.ptr(d_rgb), L.ptr(d_feat), L.stream_of(g)), "nefes_fusion_bwd")
        return (None if d_rgb is None else d_rgb.reshape(s_rgb), None if d_feat is None else d_feat.reshape(s_feat),
                None, None, None, None, None, None, None, None, None, None, *grads)


class _GemmMode:
    """The convolutions' GEMMs run on the SIMT fp32 GEMM (default: fp32 parity) or, with `module.gemm_tf32 = True`, on the tcgen05
    tf32 GEMM (operands rounded to tf32: ~3e-4 of scale)."""

    def __init__(self, tf32):
        self.tf32, self.prev = bool(tf32), 0

    def __enter__(self):
        if self.tf32:
            self.prev = L.lib().nefes_gemm_mode(1)

    def __exit__(self, *a):
        if self.tf32:
            L.lib().nefes_gemm_mode(self.prev)


def fusion_net(module, rgb, feat, B, H, W):
    """FusionNet (nerfh_nff.py:356-418) on pixel-major inputs: rgb [B*H*W,3], feat [B*H*W,128] -> [B*H*W,128].
    `module` is the torch FusionNet holding the parameters (net.0/2/4/6, net.7 = BatchNorm2d unless no_BN)."""
    convs = [module.net[i] for i in (0, 2, 4, 6)]
    params = [c.weight for c in convs] + [c.bias for c in convs]
    run_mean = run_var = None
    momentum, eps = 0.1, 1e-5
    training = module.training
    if not module.no_BN:
        bn = module.net[7]
        params += [bn.weight, bn.bias]
        run_mean, run_var, momentum, eps = bn.running_mean, bn.running_var, bn.momentum, bn.eps
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    return _FusionNetFn.apply(rgb, feat, int(B), int(H), int(W), training, bool(module.fusion_residule), momentum, eps, run_mean, run_var,
                              bool(getattr(module, "gemm_tf32", False)), *params)


class _AffineColorFn(Function):
    @staticmethod
    def forward(ctx, rgb, hist, params, B):
        L.need_cuda(rgb, hist, params)
        rgb_c, hist_c, p_c = L.f32c(rgb), L.f32c(hist), L.f32c(params)
        N = rgb_c.shape[0]
        if rgb_c.dim() != 2 or rgb_c.shape[1] != 3 or N % B or hist_c.shape != (B, 10):
            raise RuntimeError(f"nefes_b200: affine colour transform expects rgb [B*N,3] and hist [B,10], got {tuple(rgb.shape)} / {tuple(hist.shape)}")
        dev = rgb_c.device
        ab, hid, out = torch.empty(B, 12, device=dev), torch.empty(B, 96, device=dev), torch.empty(N, 3, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().nefes_affine_color_fwd(L.ptr(p_c), L.ptr(hist_c), L.ptr(rgb_c), B, N // B, L.ptr(ab), L.ptr(hid), L.ptr(out),
                                                   L.stream_of(rgb_c)), "nefes_affine_color_fwd")
        ctx.save_for_backward(rgb_c, hist_c, p_c, ab, hid, out)
        ctx.B = B
        ctx.mark_non_differentiable(ab)
        return out, ab

    @staticmethod
    def backward(ctx, g, _g_ab):
        rgb_c, hist_c, p_c, ab, hid, out = ctx.saved_tensors
        B, N = ctx.B, rgb_c.shape[0]
        dev = rgb_c.device
        g = L.f32c(g)
        d_ab = torch.empty(B, 12, device=dev)
        d_rgb = torch.empty(N, 3, device=dev) if ctx.needs_input_grad[0] else None
        d_p = torch.zeros_like(p_c) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(dev):
            L.check(L.lib().nefes_affine_color_bwd(L.ptr(p_c), L.ptr(hist_c), L.ptr(rgb_c), L.ptr(out), L.ptr(g), L.ptr(ab), L.ptr(hid), B, N // B,
                                                   L.ptr(d_ab), L.ptr(d_rgb), L.ptr(d_p), L.stream_of(g)), "nefes_affine_color_bwd")
        return d_rgb, None, d_p, None


def affine_color(exposure_params, rgb, hist, B):
    """nerfh_nff.py:605-626 on the engine: rgb [B*N,3], hist [B,10] -> (sigmoid(K_b rgb + b_b) [B*N,3], [K_b | b_b] [B,12])."""
    return _AffineColorFn.apply(rgb, hist, exposure_params, int(B))
