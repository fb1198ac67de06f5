"""ctypes binding of libnefes_b200.so (the C-ABI declared in include/nefes_b200.h).

This is the only place the shared library is touched.  There is no fallback: if the library is
missing or a call fails, a RuntimeError is raised (SURVEY.md section 8b "Errors").
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnefes_b200.so")

MAX_LAYERS = 18
NET_COARSE, NET_FINE = 0, 1
MODE_SIGMA, MODE_STATIC, MODE_FULL = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_TF32 = 0, 1, 2
COMP_SIGMA, COMP_STATIC, COMP_TRANSIENT, COMP_TRANSIENT_STATIC_ONLY = 0, 1, 2, 3
RAW_CH = {MODE_SIGMA: 1, MODE_STATIC: 132, MODE_FULL: 137}

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float


class Layout(C.Structure):
    _fields_ = [("n_layers", i32), ("n_params", i64), ("out_dim", i32 * MAX_LAYERS),
                ("in_dim", i32 * MAX_LAYERS), ("w_off", i64 * MAX_LAYERS), ("b_off", i64 * MAX_LAYERS),
                ("name", C.c_char_p * MAX_LAYERS)]


class CompOut(C.Structure):
    _fields_ = [(n, vp) for n in ("rgb", "feat", "disp", "acc", "weights", "depth", "beta", "tsig")]


class CompGrad(C.Structure):
    _fields_ = [(n, vp) for n in ("rgb", "feat", "disp", "acc", "weights", "depth", "beta", "tsig")]


class RenderCfg(C.Structure):
    _fields_ = [(n, i32) for n in ("n_samples", "n_importance", "prec", "test_time", "output_transient",
                                   "transient_at_test", "net_coarse", "net_fine")] + [("beta_min", f32), ("forward_only", i32), ("weights_packed", i32)]


class RenderIn(C.Structure):
    _fields_ = [("rays", vp), ("ld_rays", i32), ("params_coarse", vp), ("params_fine", vp), ("t_vals", vp), ("t_rand", vp),
                ("u", vp), ("u_per_ray", i32), ("noise_coarse", vp), ("noise_fine", vp)]


class RenderOut(C.Structure):
    _fields_ = [("coarse", CompOut), ("fine", CompOut), ("z_coarse", vp), ("z_fine", vp), ("z_samples", vp), ("inds", vp),
                ("z_std", vp)]


class HashLevel(C.Structure):
    _fields_ = [("scale", f32), ("res", C.c_uint32), ("size", C.c_uint32), ("offset", C.c_uint32), ("dense", C.c_uint32)]


class HashLayout(C.Structure):
    _fields_ = [("n_levels", i32), ("n_entries", i64), ("level", HashLevel * 32)]


_SIGS = {
    "nefes_version": (i32, []),
    "nefes_last_error": (C.c_char_p, []),
    "nefes_launch_count": (i64, []),
    "nefes_param_layout": (i32, [i32, C.POINTER(Layout)]),
    "nefes_get_rays_fwd": (i32, [vp, i32, i32, i32, f32, vp, vp, vp]),
    "nefes_get_rays_bwd": (i32, [vp, vp, i32, i32, i32, f32, vp, vp]),
    "nefes_sample_coarse": (i32, [vp, vp, i32, vp, vp, i32, i32, vp, vp]),
    "nefes_sample_pdf": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "nefes_sample_pdf_from_cdf": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    "nefes_sample_fine": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "nefes_encode_pe_fwd": (i32, [vp, i64, i32, vp, i32, vp]),
    "nefes_encode_pe_bwd": (i32, [vp, vp, i32, i64, i32, vp, vp]),
    "nefes_hash_layout": (i32, [i32, i32, i32, f32, C.POINTER(HashLayout)]),
    "nefes_encode_hash_fwd": (i32, [vp, vp, i64, C.POINTER(HashLayout), vp, vp]),
    "nefes_encode_hash_bwd": (i32, [vp, vp, vp, i64, C.POINTER(HashLayout), vp, vp, vp]),
    "nefes_encode_sh_fwd": (i32, [vp, i64, vp, vp]),
    "nefes_encode_sh_bwd": (i32, [vp, vp, i64, vp, vp]),
    "nefes_linear_fwd": (i32, [vp, i64, vp, vp, vp, i64, i64, i32, i32, i32, vp]),
    "nefes_linear_dgrad": (i32, [vp, i64, vp, vp, i64, i64, i32, i32, vp, i64, vp]),
    "nefes_linear_wgrad": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, vp]),
    "nefes_mlp_workspace": (i32, [i32, i32, i32, i64, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "nefes_mlp_fwd": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp]),
    "nefes_mlp_bwd": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "nefes_mlp_fwd_tiles": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp]),
    "nefes_mlp_bwd_tiles": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "nefes_composite_fwd": (i32, [vp, vp, vp, i32, i32, i32, f32, C.POINTER(CompOut), vp]),
    "nefes_composite_bwd": (i32, [vp, vp, vp, i32, i32, i32, C.POINTER(CompGrad), vp, vp]),
    "nefes_composite_fwd_tiles": (i32, [vp, vp, vp, i32, i32, i32, f32, C.POINTER(CompOut), vp]),
    "nefes_composite_bwd_tiles": (i32, [vp, vp, vp, i32, i32, i32, C.POINTER(CompGrad), vp, vp]),
    "nefes_composite_bwd_compact": (i32, [vp, vp, vp, i32, i32, i32, C.POINTER(CompGrad), vp, vp]),
    "nefes_mlp_bwd_compact": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "nefes_adam_step_dev": (i32, [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, vp]),
    "nefes_nerfw_loss_fwd": (i32, [vp, vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp]),
    "nefes_nerfw_loss_bwd": (i32, [vp, vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp, vp, vp]),
    "nefes_gemm_mode": (i32, [i32]),
    "nefes_feat_loss_fwd": (i32, [vp, vp, vp, i64, i32, vp, vp, vp]),
    "nefes_feat_loss_bwd": (i32, [vp, vp, vp, vp, i64, i32, vp, vp, vp]),
    "nefes_fusion_workspace": (i64, [i64]),
    "nefes_fusion_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, vp, vp, vp]),
    "nefes_fusion_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
    "nefes_affine_color_fwd": (i32, [vp, vp, vp, i32, i64, vp, vp, vp, vp]),
    "nefes_affine_color_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp]),
    "nefes_render_rays_workspace": (i32, [C.POINTER(RenderCfg), i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "nefes_render_rays_prepack": (i32, [C.POINTER(RenderCfg), C.POINTER(RenderIn), i64, vp, vp]),
    "nefes_render_rays_fwd": (i32, [C.POINTER(RenderCfg), C.POINTER(RenderIn), i64, C.POINTER(RenderOut), vp, vp, vp]),
    "nefes_render_rays_bwd": (i32, [C.POINTER(RenderCfg), C.POINTER(RenderIn), i64, C.POINTER(RenderOut), C.POINTER(CompGrad),
                                    C.POINTER(CompGrad), vp, vp, vp, vp, vp, vp]),
    "nefes_mlp_dgrad": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]),
    "nefes_mlp_wgrad": (i32, [vp, i32, i32, i32, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp]),
    "nefes_pose_rays_fwd": (i32, [vp, vp, i32, i32, f32, f32, f32, vp, vp, i32, vp, vp]),
    "nefes_pose_rays_bwd": (i32, [vp, vp, i32, i32, i32, f32, vp, vp]),
    "nefes_cosine_loss_fwd": (i32, [vp, vp, vp, i32, i32, vp, vp]),
    "nefes_cosine_loss_bwd": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, vp, vp]),
    "nefes_upsample_crop_fwd": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "nefes_upsample_crop_bwd": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "nefes_pose_adam_step": (i32, [vp, vp, vp, vp, i32, vp, f32, f32, f32, f32, f32, vp, vp]),
    "nefes_prof_enable": (i32, [i32]),
    "nefes_prof_report": (i32, [C.c_char_p, i32]),
    "nefes_adam_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"nefes_b200: {LIB_PATH} is missing. Build it with `make` at the repo root "
                "(or __graft_entry__.build()). There is no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)      # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(code: int, what: str):
    if code != 0:
        msg = lib().nefes_last_error()
        raise RuntimeError(f"nefes_b200: {what} failed (code {code}): {msg.decode() if msg else '?'}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_of(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nefes_b200: tensors must live on a CUDA device (no CPU fallback)")


def f32c(t):
    """contiguous fp32 view/copy"""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_layouts = {}


def layout(net: int):
    """Flat parameter layout of net (0 coarse, 1 fine): list of (name, out, in, w_off, b_off), n_params."""
    if net not in _layouts:
        L = Layout()
        check(lib().nefes_param_layout(net, C.byref(L)), "nefes_param_layout")
        rows = [(L.name[i].decode(), L.out_dim[i], L.in_dim[i], L.w_off[i], L.b_off[i]) for i in range(L.n_layers)]
        _layouts[net] = (rows, int(L.n_params))
    return _layouts[net]


def mlp_workspace(net, mode, prec, M, N):
    a, b, c = i64(), i64(), i64()
    check(lib().nefes_mlp_workspace(net, mode, prec, M, N, C.byref(a), C.byref(b), C.byref(c)), "nefes_mlp_workspace")
    return a.value, b.value, c.value
