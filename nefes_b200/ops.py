"""torch.autograd wrappers around the C-ABI kernels.  One Function per stage of the path so each
can stand in for the matching reference callable on its own."""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function

from . import _lib as L

_lin_cache = {}

# bench.py sets this to a list to collect (tag, start_event, end_event) around every field-MLP launch
# group on the launching stream (roofline: achieved = algorithmic FLOPs / measured duration).
PROFILE = None


class _Timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if PROFILE is not None:
            self.s, self.e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e.record()
            PROFILE.append((self.tag, self.s, self.e))


def linspace01(n: int, device):
    """torch.linspace(0,1,n) as the reference builds it (rendering.py:94, :33) -- produced by
    torch on the host so the fp32 knots are bit-identical, then kept on the device."""
    key = (n, str(device))
    if key not in _lin_cache:
        _lin_cache[key] = torch.linspace(0., 1., steps=n).to(device)
    return _lin_cache[key]


def _buf(nbytes: int, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# K1 get_rays
# ------------------------------------------------------------------------------------------------
class _GetRays(Function):
    @staticmethod
    def forward(ctx, c2w, H, W, focal):
        L.need_cuda(c2w)
        m = L.f32c(c2w[..., :3, :4])
        B = m.shape[0]
        o = torch.empty(B, H, W, 3, device=m.device)
        d = torch.empty_like(o)
        L.check(L.lib().nefes_get_rays_fwd(L.ptr(m), B, H, W, float(focal), L.ptr(o), L.ptr(d), L.stream_of(m)),
                "nefes_get_rays_fwd")
        ctx.meta = (B, H, W, float(focal), tuple(c2w.shape))
        return o, d

    @staticmethod
    def backward(ctx, g_o, g_d):
        B, H, W, focal, shape = ctx.meta
        g_o, g_d = L.f32c(g_o), L.f32c(g_d)
        ref = g_o if g_o is not None else g_d
        out = torch.empty(B, 3, 4, device=ref.device)
        L.check(L.lib().nefes_get_rays_bwd(L.ptr(g_o), L.ptr(g_d), B, H, W, focal, L.ptr(out), L.stream_of(ref)),
                "nefes_get_rays_bwd")
        if shape[-2] == 3 and shape[-1] == 4:
            return out.reshape(shape), None, None, None
        full = torch.zeros(shape, device=ref.device)
        full[..., :3, :4] = out.reshape(shape[:-2] + (3, 4))
        return full, None, None, None


def get_rays(H, W, focal, c2w_batched):
    """c2w [B,3|4,4] -> rays_o, rays_d [B,H,W,3]"""
    return _GetRays.apply(c2w_batched, int(H), int(W), float(focal))


# ------------------------------------------------------------------------------------------------
# K2 / K3 sampling (no gradients flow through sample positions: rendering.py:51,136)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def sample_coarse(near, far, ld, n_rays, n_samples, t_rand=None):
    """near/far: fp32 CUDA tensors addressed as near[i*ld]."""
    L.need_cuda(near, far, t_rand)
    t_vals = linspace01(n_samples, near.device)
    z = torch.empty(n_rays, n_samples, device=near.device)
    t_rand = L.f32c(t_rand)
    L.check(L.lib().nefes_sample_coarse(L.ptr(near), L.ptr(far), ld, L.ptr(t_vals), L.ptr(t_rand), n_rays,
                                        n_samples, L.ptr(z), L.stream_of(near)), "nefes_sample_coarse")
    return z


@torch.no_grad()
def sample_pdf(bins, weights, n_samples, u=None, cdf=None, return_inds=False):
    """u None -> det=True (linspace).  cdf given -> skip the pdf->cdf step."""
    L.need_cuda(bins, weights, u, cdf)
    bins = L.f32c(bins)
    N, nb = bins.shape
    per_ray = 1
    if u is None:
        u, per_ray = linspace01(n_samples, bins.device), 0
    u = L.f32c(u)
    out = torch.empty(N, n_samples, device=bins.device)
    inds = torch.empty(N, n_samples, dtype=torch.int32, device=bins.device) if return_inds else None
    st = L.stream_of(bins)
    if cdf is not None:
        cdf = L.f32c(cdf)
        L.check(L.lib().nefes_sample_pdf_from_cdf(L.ptr(bins), L.ptr(cdf), L.ptr(u), per_ray, N, nb, n_samples,
                                                  L.ptr(out), L.ptr(inds), st), "nefes_sample_pdf_from_cdf")
    else:
        weights = L.f32c(weights)
        L.check(L.lib().nefes_sample_pdf(L.ptr(bins), L.ptr(weights), L.ptr(u), per_ray, N, nb, n_samples,
                                         L.ptr(out), L.ptr(inds), None, st), "nefes_sample_pdf")
    return (out, inds) if return_inds else out


@torch.no_grad()
def sample_fine(z_coarse, weights_coarse, n_fine, u=None, want_aux=True):
    """mids -> sample_pdf(weights[:,1:-1]) -> sort(cat) in one kernel.  Returns z_fine, z_samples, inds."""
    L.need_cuda(z_coarse, weights_coarse, u)
    z_coarse, weights_coarse = L.f32c(z_coarse), L.f32c(weights_coarse)
    N, S = z_coarse.shape
    per_ray = 1
    if u is None:
        u, per_ray = linspace01(n_fine, z_coarse.device), 0
    u = L.f32c(u)
    z_fine = torch.empty(N, S + n_fine, device=z_coarse.device)
    z_s = torch.empty(N, n_fine, device=z_coarse.device) if want_aux else None
    inds = torch.empty(N, n_fine, dtype=torch.int32, device=z_coarse.device) if want_aux else None
    L.check(L.lib().nefes_sample_fine(L.ptr(z_coarse), L.ptr(weights_coarse), L.ptr(u), per_ray, N, S, n_fine,
                                      L.ptr(z_fine), L.ptr(z_s), L.ptr(inds), L.stream_of(z_coarse)),
            "nefes_sample_fine")
    return z_fine, z_s, inds


# ------------------------------------------------------------------------------------------------
# K4a positional encoding
# ------------------------------------------------------------------------------------------------
class _EncodePE(Function):
    @staticmethod
    def forward(ctx, x, n_freqs):
        L.need_cuda(x)
        xc = L.f32c(x.reshape(-1, 3))
        M = xc.shape[0]
        ch = 3 + 6 * n_freqs
        out = torch.empty(M, ch, device=xc.device)
        L.check(L.lib().nefes_encode_pe_fwd(L.ptr(xc), M, n_freqs, L.ptr(out), ch, L.stream_of(xc)),
                "nefes_encode_pe_fwd")
        ctx.save_for_backward(xc)
        ctx.meta = (n_freqs, tuple(x.shape))
        return out.reshape(tuple(x.shape[:-1]) + (ch,))

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        n_freqs, shape = ctx.meta
        ch = 3 + 6 * n_freqs
        g = L.f32c(g.reshape(-1, ch))
        dx = torch.empty_like(xc)
        L.check(L.lib().nefes_encode_pe_bwd(L.ptr(xc), L.ptr(g), ch, xc.shape[0], n_freqs, L.ptr(dx),
                                            L.stream_of(xc)), "nefes_encode_pe_bwd")
        return dx.reshape(shape), None


def encode_pe(x, n_freqs):
    return _EncodePE.apply(x, int(n_freqs))


# ------------------------------------------------------------------------------------------------
# K5 field query (PE + MLP)
# ------------------------------------------------------------------------------------------------
class TiledRaw:
    """raw [N,S,C] held in the engine's tile-major layout: `t` is [ceil(N*S/128), C, 128] fp32 (blocks of 128
    consecutive points, channel-major inside a block -- what the fused MLP chain writes with coalesced stores and the
    compositing kernels read).  Stays inside render_rays; `.rows()` gives the reference's [N,S,C] tensor."""

    def __init__(self, t, N, S, C_):
        self.t, self.N, self.S, self.C = t, N, S, C_
        # set by render_rays, the only consumer of its own raw: the compositing backward may then hand the field backward
        # the COMPACT cotangent (weights x per-ray cotangents) instead of a [T,C,128] fp32 block
        self.private = False
        self.single = True                           # produced by ONE field query (its autograd node owns `t`)

    @property
    def shape(self):
        return torch.Size((self.N, self.S, self.C))

    @property
    def device(self):
        return self.t.device

    def rows(self):
        M = self.N * self.S
        return self.t.permute(0, 2, 1).reshape(-1, self.C)[:M].reshape(self.N, self.S, self.C)

    @staticmethod
    def cat(parts):
        if any((p.N * p.S) % 128 for p in parts[:-1]):
            raise RuntimeError("nefes_b200: netchunk must cut the rays at multiples of 128 points")
        out = TiledRaw(torch.cat([p.t for p in parts], 0), sum(p.N for p in parts), parts[0].S, parts[0].C)
        out.single = False
        return out


_tiled_depth = 0
# compact cotangents in flight between _Composite.backward and _FieldQuery.backward, keyed by raw's storage address
_COMPACT = {}


class tiled_raw:
    """Context manager used by render_rays: field queries issued inside return TiledRaw where the engine can
    (bf16 path, samples per ray dividing 128), so raw never takes the row-major detour between the MLP and the
    compositing kernels."""

    def __enter__(self):
        global _tiled_depth
        _tiled_depth += 1

    def __exit__(self, *a):
        global _tiled_depth
        _tiled_depth -= 1


def want_tiled(prec, mode, S):
    return _tiled_depth > 0 and prec == L.PREC_BF16 and mode != L.MODE_SIGMA and 128 % int(S) == 0


class _FieldQuery(Function):
    """pts [N,S,3], dirs [N,3] (or None for MODE_SIGMA), flat params -> raw [N,S,C]
    (tiled=True: the tile-major block tensor [T,C,128] of TiledRaw)."""

    @staticmethod
    def forward(ctx, pts, dirs, flat, net, mode, prec, tiled, grad_on=True):
        L.need_cuda(pts, dirs, flat)
        pts_c = L.f32c(pts)
        N, S = pts_c.shape[0], pts_c.shape[1]
        dirs_c = L.f32c(dirs) if dirs is not None else None
        flat_c = flat.detach()
        if not flat_c.is_contiguous() or flat_c.dtype != torch.float32:
            raise RuntimeError("nefes_b200: flat parameter buffer must be contiguous fp32")
        Cc = L.RAW_CH[mode]
        dev = pts_c.device
        # needs_input_grad reports requires_grad of the inputs even under torch.no_grad(); grad_on is
        # torch.is_grad_enabled() sampled by the caller (inside forward() it is always False)
        need_bwd = bool(grad_on) and any(ctx.needs_input_grad[:3])
        sv, sf, _ = L.mlp_workspace(net, mode, prec, N * S, N)
        saved = _buf(sv, dev)
        scratch = _buf(sf, dev)
        raw = torch.empty((N * S + 127) // 128, Cc, 128, device=dev) if tiled else torch.empty(N, S, Cc, device=dev)
        fn = L.lib().nefes_mlp_fwd_tiles if tiled else L.lib().nefes_mlp_fwd
        with torch.cuda.device(dev), _Timed("mlp_fwd"):
            L.check(fn(L.ptr(flat_c), net, mode, prec, L.ptr(pts_c), L.ptr(dirs_c), N, S,
                       L.ptr(raw), L.ptr(saved), L.ptr(scratch), L.stream_of(pts_c)), "nefes_mlp_fwd")
        if need_bwd:
            ctx.save_for_backward(pts_c, dirs_c, flat_c, raw, saved)
            ctx.meta = (net, mode, prec, N, S, tuple(pts.shape), None if dirs is None else tuple(dirs.shape), tiled)
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        pts_c, dirs_c, flat_c, raw, saved = ctx.saved_tensors
        net, mode, prec, N, S, pshape, dshape, tiled = ctx.meta
        dev = pts_c.device
        compact = _COMPACT.pop(raw.data_ptr(), None) if tiled else None
        if compact is None:
            if tiled and d_raw is not None and d_raw.numel() > 1 and all(s_ == 0 for s_ in d_raw.stride()):
                raise RuntimeError("nefes_b200: the compact cotangent of a private tile-major raw is gone (the placeholder "
                                   "d_raw reached the field backward); was the graph backpropagated twice or ops._COMPACT cleared "
                                   "mid-backward?")
            d_raw = L.f32c(d_raw)
        need_p, need_d, need_w = ctx.needs_input_grad[:3]
        d_pts = torch.empty_like(pts_c) if need_p else None
        d_dirs = torch.empty_like(dirs_c) if (need_d and dirs_c is not None) else None
        d_flat = torch.zeros_like(flat_c) if need_w else None
        _, _, sb = L.mlp_workspace(net, mode, prec, N * S, N)
        scratch = _buf(sb, dev)
        fn = L.lib().nefes_mlp_bwd_tiles if tiled else L.lib().nefes_mlp_bwd
        with torch.cuda.device(dev), _Timed("mlp_bwd"):
            if compact is not None:
                cg, g_rgb, g_feat = compact
                L.check(L.lib().nefes_mlp_bwd_compact(L.ptr(flat_c), net, mode, prec, L.ptr(pts_c), L.ptr(dirs_c), N, S,
                                                      L.ptr(raw), L.ptr(cg), L.ptr(g_rgb), L.ptr(g_feat), L.ptr(saved),
                                                      L.ptr(scratch), L.ptr(d_flat), L.ptr(d_pts), L.ptr(d_dirs),
                                                      L.stream_of(pts_c)), "nefes_mlp_bwd_compact")
            else:
                L.check(fn(L.ptr(flat_c), net, mode, prec, L.ptr(pts_c), L.ptr(dirs_c), N, S,
                           L.ptr(raw), L.ptr(d_raw), L.ptr(saved), L.ptr(scratch), L.ptr(d_flat),
                           L.ptr(d_pts), L.ptr(d_dirs), L.stream_of(pts_c)), "nefes_mlp_bwd")
        return (d_pts.reshape(pshape) if d_pts is not None else None,
                d_dirs.reshape(dshape) if d_dirs is not None else None, d_flat, None, None, None, None, None)


def field_query(pts, dirs, flat, net, mode, prec=L.PREC_FP32):
    """-> raw [N,S,C]; inside a `tiled_raw()` block (render_rays) a TiledRaw where the engine supports it."""
    grad_on = torch.is_grad_enabled()
    if want_tiled(prec, mode, pts.shape[1]):
        t = _FieldQuery.apply(pts, dirs, flat, int(net), int(mode), int(prec), True, grad_on)
        return TiledRaw(t, pts.shape[0], pts.shape[1], L.RAW_CH[mode])
    return _FieldQuery.apply(pts, dirs, flat, int(net), int(mode), int(prec), False, grad_on)


# ------------------------------------------------------------------------------------------------
# K6 compositing
# ------------------------------------------------------------------------------------------------
class _Composite(Function):
    """raw [N,S,C], z [N,S] -> rgb, feat, disp, acc, weights, depth, beta[, transient_sigmas].
    transient_sigmas (= raw[...,135]) is an output of the op so its cotangent reaches d_raw inside the
    backward kernel instead of through a dense zero-filled slice gradient."""

    @staticmethod
    def forward(ctx, raw, z, noise, mode, beta_min, tiled, compact=False):
        ctx.set_materialize_grads(False)             # unused outputs arrive as None, not as zero-filled tensors
        L.need_cuda(raw, z, noise)
        raw_c, z_c, noise_c = L.f32c(raw), L.f32c(z), L.f32c(noise)
        N, S = z_c.shape
        dev = raw_c.device
        acc = torch.empty(N, device=dev)
        weights = torch.empty(N, S, device=dev)
        tsig = None
        if mode == L.COMP_SIGMA:
            rgb = feat = disp = depth = beta = None
        else:
            rgb = torch.empty(N, 3, device=dev)
            feat = torch.empty(N, 128, device=dev)
            disp, depth, beta = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
            if mode in (L.COMP_TRANSIENT, L.COMP_TRANSIENT_STATIC_ONLY):
                tsig = torch.empty(N, S, device=dev)
        out = L.CompOut(L.ptr(rgb), L.ptr(feat), L.ptr(disp), L.ptr(acc), L.ptr(weights), L.ptr(depth), L.ptr(beta),
                        L.ptr(tsig))
        fn = L.lib().nefes_composite_fwd_tiles if tiled else L.lib().nefes_composite_fwd
        with torch.cuda.device(dev):
            L.check(fn(L.ptr(raw_c), L.ptr(z_c), L.ptr(noise_c), N, S, mode,
                       float(beta_min), C.byref(out), L.stream_of(raw_c)), "nefes_composite_fwd")
        ctx.save_for_backward(raw_c, z_c, noise_c)
        ctx.meta = (mode, N, S, tuple(raw.shape), tiled, bool(compact) and tiled and mode != L.COMP_SIGMA)
        if mode == L.COMP_SIGMA:
            return acc, weights
        if tsig is None:
            return rgb, feat, disp, acc, weights, depth, beta
        return rgb, feat, disp, acc, weights, depth, beta, tsig

    @staticmethod
    def backward(ctx, *grads):
        raw_c, z_c, noise_c = ctx.saved_tensors
        mode, N, S, rshape, tiled, compact = ctx.meta
        if mode == L.COMP_SIGMA:
            g_acc, g_w = grads
            g = dict(acc=g_acc, weights=g_w)
        else:
            g = dict(zip(("rgb", "feat", "disp", "acc", "weights", "depth", "beta", "tsig"), grads))
        g = {k: L.f32c(v) for k, v in g.items() if v is not None}
        gs = L.CompGrad(*[L.ptr(g.get(k)) for k in ("rgb", "feat", "disp", "acc", "weights", "depth", "beta", "tsig")])
        if not ctx.needs_input_grad[0]:              # raw is a constant here (frozen field, no gradient to the rays)
            return None, None, None, None, None, None, None
        if compact:
            # the field backward picks the compact cotangent up by raw's address; what autograd carries is a placeholder
            # of the right shape that is never read (a 0-dim tensor expanded, no memory, no kernel)
            cg = torch.empty(N, 5, S, device=raw_c.device)
            with torch.cuda.device(raw_c.device):
                L.check(L.lib().nefes_composite_bwd_compact(L.ptr(raw_c), L.ptr(z_c), L.ptr(noise_c), N, S, mode, C.byref(gs),
                                                            L.ptr(cg), L.stream_of(raw_c)), "nefes_composite_bwd_compact")
            while len(_COMPACT) >= 4096:             # leftovers of aborted backward passes: drop the oldest
                _COMPACT.pop(next(iter(_COMPACT)))
            _COMPACT[raw_c.data_ptr()] = (cg, g.get("rgb"), g.get("feat"))
            return cg.new_zeros(()).expand(rshape), None, None, None, None, None, None
        d_raw = torch.empty_like(raw_c)
        fn = L.lib().nefes_composite_bwd_tiles if tiled else L.lib().nefes_composite_bwd
        with torch.cuda.device(raw_c.device):
            L.check(fn(L.ptr(raw_c), L.ptr(z_c), L.ptr(noise_c), N, S, mode, C.byref(gs),
                       L.ptr(d_raw), L.stream_of(raw_c)), "nefes_composite_bwd")
        return d_raw.reshape(rshape), None, None, None, None, None, None


def composite(raw, z, noise, mode, beta_min=0.1):
    if isinstance(raw, TiledRaw):
        return _Composite.apply(raw.t, z, noise, int(mode), float(beta_min), True, raw.private)
    return _Composite.apply(raw, z, noise, int(mode), float(beta_min), False, False)


# ------------------------------------------------------------------------------------------------
# the whole path as one engine call (nefes_render_rays_fwd / _bwd)
# ------------------------------------------------------------------------------------------------
class _RenderRays(Function):
    """ray_batch [N, >=11], flat_coarse, flat_fine (+ explicit random draws) -> the composited outputs of both passes.
    One autograd node for rendering.py:68-180: sample points, raw, saved activations and compact cotangents never leave
    the engine's two workspaces."""

    @staticmethod
    def forward(ctx, rays, flat_c, flat_f, t_rand, u, noise_c, noise_f, cfg):
        ctx.set_materialize_grads(False)
        L.need_cuda(rays, flat_c, flat_f, t_rand, u, noise_c, noise_f)
        rays_c = L.f32c(rays)
        N, ld = rays_c.shape
        dev = rays_c.device
        S, ni = cfg["n_samples"], cfg["n_importance"]
        Sf = S + ni
        need_bwd = bool(cfg.get("grad_on", True)) and any(ctx.needs_input_grad[:3])
        c = L.RenderCfg(S, ni, cfg["prec"], int(cfg["test_time"]), int(cfg["output_transient"]), int(cfg["transient_at_test"]),
                        cfg["net_coarse"], cfg["net_fine"], float(cfg["beta_min"]), 0 if need_bwd else 1)
        kb, sfb, sbb = C.c_int64(), C.c_int64(), C.c_int64()
        L.check(L.lib().nefes_render_rays_workspace(C.byref(c), N, C.byref(kb), C.byref(sfb), C.byref(sbb)),
                "nefes_render_rays_workspace")
        keep, scratch = _buf(kb.value, dev), _buf(sfb.value, dev)
        t_rand, u, noise_c, noise_f = L.f32c(t_rand), L.f32c(u), L.f32c(noise_c), L.f32c(noise_f)
        per_ray = 1
        if u is None:
            u, per_ray = linspace01(ni, dev), 0
        fc, ff = flat_c.detach(), flat_f.detach()
        if not (fc.is_contiguous() and ff.is_contiguous() and fc.dtype == ff.dtype == torch.float32):
            raise RuntimeError("nefes_b200: flat parameter buffers must be contiguous fp32")
        inp = L.RenderIn(L.ptr(rays_c), ld, L.ptr(fc), L.ptr(ff), L.ptr(linspace01(S, dev)), L.ptr(t_rand), L.ptr(u), per_ray,
                         L.ptr(noise_c), L.ptr(noise_f))
        e = lambda *shape: torch.empty(*shape, device=dev)
        sigma_only = bool(cfg["test_time"])
        transient = bool(cfg["output_transient"])
        acc0, w0 = e(N), e(N, S)
        rgb0 = feat0 = disp0 = depth0 = beta0 = None
        if not sigma_only:
            rgb0, feat0, disp0, depth0, beta0 = e(N, 3), e(N, 128), e(N), e(N), e(N)
        rgb, feat, disp, acc, w, depth, beta = e(N, 3), e(N, 128), e(N), e(N), e(N, Sf), e(N), e(N)
        tsig = e(N, Sf) if transient else None
        z_c, z_f, z_s, z_std = e(N, S), e(N, Sf), e(N, ni), e(N)
        inds = torch.empty(N, ni, dtype=torch.int32, device=dev)
        out = L.RenderOut(L.CompOut(L.ptr(rgb0), L.ptr(feat0), L.ptr(disp0), L.ptr(acc0), L.ptr(w0), L.ptr(depth0), L.ptr(beta0), None),
                          L.CompOut(L.ptr(rgb), L.ptr(feat), L.ptr(disp), L.ptr(acc), L.ptr(w), L.ptr(depth), L.ptr(beta), L.ptr(tsig)),
                          L.ptr(z_c), L.ptr(z_f), L.ptr(z_s), L.ptr(inds), L.ptr(z_std))
        with torch.cuda.device(dev), _Timed("render_fwd"):
            L.check(L.lib().nefes_render_rays_fwd(C.byref(c), C.byref(inp), N, C.byref(out), L.ptr(keep), L.ptr(scratch),
                                                  L.stream_of(rays_c)), "nefes_render_rays_fwd")
        if need_bwd:
            ctx.save_for_backward(rays_c, fc, ff, t_rand, u, noise_c, noise_f, z_c, z_f, keep)
            ctx.meta = (dict(cfg), per_ray, sbb.value, tuple(rays.shape))
        ctx.mark_non_differentiable(z_c, z_f, z_s, inds, z_std)
        return (rgb, feat, disp, acc, w, depth, beta, tsig, rgb0, feat0, disp0, acc0, w0, depth0, z_std, z_c, z_f, z_s, inds)

    @staticmethod
    def backward(ctx, *g):
        rays_c, fc, ff, t_rand, u, noise_c, noise_f, z_c, z_f, keep = ctx.saved_tensors
        cfg, per_ray, sbb, rshape = ctx.meta
        N, ld = rays_c.shape
        dev = rays_c.device
        S, ni = cfg["n_samples"], cfg["n_importance"]
        c = L.RenderCfg(S, ni, cfg["prec"], int(cfg["test_time"]), int(cfg["output_transient"]), int(cfg["transient_at_test"]),
                        cfg["net_coarse"], cfg["net_fine"], float(cfg["beta_min"]), 0)
        inp = L.RenderIn(L.ptr(rays_c), ld, L.ptr(fc), L.ptr(ff), L.ptr(linspace01(S, dev)), L.ptr(t_rand), L.ptr(u), per_ray,
                         L.ptr(noise_c), L.ptr(noise_f))
        nul = L.CompOut(*([None] * 8))
        out = L.RenderOut(nul, nul, L.ptr(z_c), L.ptr(z_f), None, None, None)
        gf = [L.f32c(t) for t in g[0:8]]
        gc = [L.f32c(t) for t in g[8:14]] + [None, None]
        g_fine, g_coarse = L.CompGrad(*[L.ptr(t) for t in gf]), L.CompGrad(*[L.ptr(t) for t in gc])
        need_r, need_c, need_f = ctx.needs_input_grad[:3]
        d_rays = torch.empty_like(rays_c) if need_r else None
        d_c = torch.zeros_like(fc) if need_c else None
        d_f = torch.zeros_like(ff) if need_f else None
        scratch = _buf(sbb, dev)
        with torch.cuda.device(dev), _Timed("render_bwd"):
            L.check(L.lib().nefes_render_rays_bwd(C.byref(c), C.byref(inp), N, C.byref(out), C.byref(g_coarse), C.byref(g_fine),
                                                  L.ptr(keep), L.ptr(scratch), L.ptr(d_c), L.ptr(d_f), L.ptr(d_rays),
                                                  L.stream_of(rays_c)), "nefes_render_rays_bwd")
        return (d_rays.reshape(rshape) if d_rays is not None else None, d_c, d_f, None, None, None, None, None)


def render_rays_fused(rays, flat_c, flat_f, cfg, t_rand=None, u=None, noise_c=None, noise_f=None):
    cfg = dict(cfg, grad_on=torch.is_grad_enabled())     # no_grad renders are forward-only: no saved activations
    return _RenderRays.apply(rays, flat_c, flat_f, t_rand, u, noise_c, noise_f, cfg)


class RenderCall:
    """nefes_render_rays_fwd / _bwd on buffers allocated ONCE (no autograd): the form a captured refinement iteration
    replays.  `forward()` renders the rays in `self.rays`; `backward(g_feat=..., g_rgb=...)` turns cotangents of the fine
    composited outputs into `self.d_rays` (frozen fields: no parameter gradients).  With `frozen_weights=True` the forward
    does not re-pack the parameters into the tensor path's operand images: the caller runs `prepack()` whenever they may
    have changed (the refiner: once per query)."""

    def __init__(self, n_rays, ld, cfg, flat_c, flat_f, device, frozen_weights=False):
        dev = torch.device(device)
        S, ni = cfg["n_samples"], cfg["n_importance"]
        Sf = S + ni
        self.N, self.ld, self.dev = n_rays, ld, dev
        self.cfg = L.RenderCfg(S, ni, cfg["prec"], int(cfg["test_time"]), int(cfg["output_transient"]),
                               int(cfg["transient_at_test"]), cfg["net_coarse"], cfg["net_fine"], float(cfg["beta_min"]), 0,
                               1 if frozen_weights else 0)
        kb, sfb, sbb = C.c_int64(), C.c_int64(), C.c_int64()
        L.check(L.lib().nefes_render_rays_workspace(C.byref(self.cfg), n_rays, C.byref(kb), C.byref(sfb), C.byref(sbb)),
                "nefes_render_rays_workspace")
        self.keep, self.scratch = _buf(kb.value, dev), _buf(max(sfb.value, sbb.value), dev)
        e = lambda *shape: torch.empty(*shape, device=dev)
        self.rays, self.d_rays = torch.zeros(n_rays, ld, device=dev), torch.zeros(n_rays, ld, device=dev)
        self.flat_c, self.flat_f = flat_c.detach(), flat_f.detach()
        self.t_vals, self.u = linspace01(S, dev), linspace01(ni, dev)
        sigma_only, transient = bool(cfg["test_time"]), bool(cfg["output_transient"])
        self.acc0, self.w0 = e(n_rays), e(n_rays, S)
        self.rgb0 = self.feat0 = self.disp0 = self.depth0 = self.beta0 = None
        if not sigma_only:
            self.rgb0, self.feat0, self.disp0, self.depth0, self.beta0 = e(n_rays, 3), e(n_rays, 128), e(n_rays), e(n_rays), e(n_rays)
        self.rgb, self.feat, self.disp, self.acc = e(n_rays, 3), e(n_rays, 128), e(n_rays), e(n_rays)
        self.w, self.depth, self.beta = e(n_rays, Sf), e(n_rays), e(n_rays)
        self.tsig = e(n_rays, Sf) if transient else None
        self.z_c, self.z_f = e(n_rays, S), e(n_rays, Sf)
        p = L.ptr
        self.inp = L.RenderIn(p(self.rays), ld, p(self.flat_c), p(self.flat_f), p(self.t_vals), None, p(self.u), 0, None, None)
        self.out = L.RenderOut(L.CompOut(p(self.rgb0), p(self.feat0), p(self.disp0), p(self.acc0), p(self.w0), p(self.depth0),
                                         p(self.beta0), None),
                               L.CompOut(p(self.rgb), p(self.feat), p(self.disp), p(self.acc), p(self.w), p(self.depth),
                                         p(self.beta), p(self.tsig)),
                               p(self.z_c), p(self.z_f), None, None, None)

    def prepack(self):
        with torch.cuda.device(self.dev):
            L.check(L.lib().nefes_render_rays_prepack(C.byref(self.cfg), C.byref(self.inp), self.N, L.ptr(self.keep),
                                                      L.stream_of(self.rays)), "nefes_render_rays_prepack")

    def forward(self):
        with torch.cuda.device(self.dev):
            L.check(L.lib().nefes_render_rays_fwd(C.byref(self.cfg), C.byref(self.inp), self.N, C.byref(self.out),
                                                  L.ptr(self.keep), L.ptr(self.scratch), L.stream_of(self.rays)),
                    "nefes_render_rays_fwd")

    def backward(self, g_feat=None, g_rgb=None):
        g_fine = L.CompGrad(L.ptr(g_rgb), L.ptr(g_feat), None, None, None, None, None, None)
        g_coarse = L.CompGrad(*([None] * 8))
        with torch.cuda.device(self.dev):
            L.check(L.lib().nefes_render_rays_bwd(C.byref(self.cfg), C.byref(self.inp), self.N, C.byref(self.out),
                                                  C.byref(g_coarse), C.byref(g_fine), L.ptr(self.keep), L.ptr(self.scratch),
                                                  None, None, L.ptr(self.d_rays), L.stream_of(self.rays)),
                    "nefes_render_rays_bwd")
        return self.d_rays


# ------------------------------------------------------------------------------------------------
# fused Adam on flat buffers
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    L.need_cuda(p, g, m, v)
    L.check(L.lib().nefes_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, beta1, beta2, eps,
                                    int(step), grad_scale, L.stream_of(p)), "nefes_adam_step")


@torch.no_grad()
def adam_step_dev(p, g, m, v, state2, beta1, beta2, eps, grad_scale=1.0):
    """Adam with (step count, lr) in the 2-float device tensor `state2`: safe inside a captured CUDA graph."""
    L.need_cuda(p, g, m, v, state2)
    L.check(L.lib().nefes_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), L.ptr(state2), beta1, beta2,
                                        eps, grad_scale, L.stream_of(p)), "nefes_adam_step_dev")
