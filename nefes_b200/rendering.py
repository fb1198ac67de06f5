"""The reference's render interface (script/models/rendering.py:23-243: sample_pdf, render_rays, batchify_rays, render)
on the B200 kernels.  This file is the DROP-IN BOUNDARY: callers use these names, keyword arguments, packing order of the
ray batch and result-dict keys, so `render()`'s argument handling and `render_rays()`'s output assembly necessarily restate
the reference's control flow (same conditions, same key names); the arithmetic behind them is the engine's.

Extra keyword arguments (all optional, default = reference behaviour):
    t_rand, u, noise : explicit random draws ([N,N_samples], [N,N_importance], [N,N_samples]) so a
                       caller (the parity tests) can reproduce the reference's CPU RNG stream; when
                       absent they are drawn on the device with torch's generator.
    return_aux       : also return z_vals / z_samples / inds in the result dict.
"""
import os

import torch

from . import _lib as L
from . import ops
from .nerfh_nff import _PREC, NeRFH_NFF, raw2outputs_NeRFH_NFF
from .ray_utils import get_rays


def sample_pdf(bins, weights, N_samples, det=False, pytest=False, u=None):
    """Drop-in for rendering.py:23.  det=True -> u = linspace(0,1,N_samples)."""
    if pytest:                                       # rendering.py:38-47: numpy's fixed stream
        import numpy as np
        np.random.seed(0)
        shape = list(bins.shape[:-1]) + [N_samples]
        u = torch.Tensor(np.broadcast_to(np.linspace(0., 1., N_samples), shape).copy() if det
                         else np.random.rand(*shape)).to(bins.device)
    elif det:
        u = None
    elif u is None:
        u = torch.rand(list(bins.shape[:-1]) + [N_samples], device=bins.device)
    lead = bins.shape[:-1]
    out = ops.sample_pdf(bins.reshape(-1, bins.shape[-1]), weights.reshape(-1, weights.shape[-1]), N_samples,
                         u=None if u is None else u.reshape(-1, N_samples))
    return out.reshape(*lead, N_samples)


# NEFES_FUSED_RENDER=0 keeps render_rays on the staged path (one autograd node per stage) even where the one-call path applies
FUSED_RENDER = os.environ.get("NEFES_FUSED_RENDER", "1") != "0"


def _fused_applies(ray_batch, network_fn, network_query_fn, N_samples, N_importance, network_fine, args):
    """The whole-path engine call covers the reference's own configuration: the standard query function, both fields on
    the engine with one arithmetic, a fine pass, view directions present.  Rays beyond one netchunk of fine points are
    handed over in netchunk-sized groups (render_rays): rays are independent, so the numbers are those of one call."""
    return (FUSED_RENDER and N_importance > 0 and getattr(network_query_fn, "nefes_standard", False)
            and isinstance(network_fn, NeRFH_NFF) and isinstance(network_fine, NeRFH_NFF)
            and network_fn.precision == network_fine.precision and not args.use_fine_only
            and ray_batch.shape[1] >= 11 and N_samples + N_importance <= 256
            and network_query_fn.netchunk >= 2 * (N_samples + N_importance))


def _render_rays_fused(ray_batch, network_fn, N_samples, perturb, N_importance, network_fine, raw_noise_std, pytest,
                       test_time, args, t_rand, u, noise, return_aux):
    N_rays, dev = ray_batch.shape[0], ray_batch.device
    output_transient = bool(args.NeRFW)
    if output_transient and network_fine.net_id != L.NET_FINE:
        raise RuntimeError("nefes_b200: transient output requested from a field without transient heads")
    if perturb > 0. and t_rand is None:
        t_rand = torch.rand(N_rays, N_samples, device=dev)                        # rendering.py:110
    if noise is None and raw_noise_std > 0. and not test_time:
        noise = torch.randn(N_rays, N_samples, device=dev) * raw_noise_std        # nerfh_nff.py:67
    det = (perturb == 0.)
    if not det and u is None:
        u = torch.rand(N_rays, N_importance, device=dev)                          # rendering.py:36
    if pytest:
        import numpy as np
        np.random.seed(0)
        u = None if det else torch.Tensor(np.random.rand(N_rays, N_importance)).to(dev)
    noise_f = None
    if raw_noise_std > 0. and not output_transient:
        noise_f = torch.randn(N_rays, N_samples + N_importance, device=dev) * raw_noise_std
    cfg = dict(n_samples=int(N_samples), n_importance=int(N_importance), prec=_PREC[network_fn.precision],
               test_time=bool(test_time), output_transient=output_transient, transient_at_test=bool(args.transient_at_test),
               net_coarse=network_fn.net_id, net_fine=network_fine.net_id, beta_min=network_fine.beta_min)
    (rgb, feat, disp, acc, w, depth, beta, tsig, rgb0, feat0, disp0, acc0, w0, depth0, z_std, z_c, z_f, z_s, inds) = \
        ops.render_rays_fused(ray_batch, network_fn.flat, network_fine.flat, cfg, t_rand=t_rand if perturb > 0. else None,
                              u=None if det else u, noise_c=noise, noise_f=noise_f)
    ret = {'rgb_map': rgb, 'disp_map': disp, 'acc_map': acc}
    if args.nerfh_nff:
        ret['feat_map'] = feat
    if not test_time:
        ret['rgb0'], ret['disp0'], ret['acc0'] = rgb0, disp0, acc0
        ret['z_std'] = z_std
        if args.NeRFW:
            ret['transient_sigmas'] = tsig
            ret['beta'] = beta
        if args.nerfh_nff and feat0 is not None:
            ret['feat0'] = feat0
    if return_aux:
        aux = dict(z_coarse=z_c, z_samples=z_s, inds=inds, z_fine=z_f, weights_coarse=w0, depth_map=depth, weights_fine=w)
        for k, v in aux.items():
            ret['aux_' + k] = v
    return ret


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                i_epoch=-1, embedding_a=None, embedding_t=None, test_time=False, args=None, volume=None,
                t_rand=None, u=None, noise=None, return_aux=False):
    """Drop-in for rendering.py:68-180.  lindisp / white_bkgd (dormant in every reference config) take the staged route."""
    L.need_cuda(ray_batch)
    ray_batch = ray_batch if ray_batch.dtype == torch.float32 else ray_batch.float()
    if not ray_batch.is_contiguous():
        ray_batch = ray_batch.contiguous()
    if not lindisp and not white_bkgd and \
            _fused_applies(ray_batch, network_fn, network_query_fn, N_samples, N_importance, network_fine, args):
        # the reference's defaults (chunk 32768 rays, netchunk 2^21 points) put 2^22 fine points into one render_rays call:
        # two engine calls of 16384 rays each (measured: inference 8.0 M rays/s this way against 5.2 M on the staged route)
        per = network_query_fn.netchunk // (N_samples + N_importance)
        per -= per % 2
        n_all = ray_batch.shape[0]
        if n_all <= per:
            return _render_rays_fused(ray_batch, network_fn, N_samples, perturb, N_importance, network_fine, raw_noise_std, pytest,
                                      test_time, args, t_rand, u, noise, return_aux)
        if pytest:
            raise RuntimeError("nefes_b200: the pytest=True determinism hook needs the rays of a call to fit one netchunk")
        cut = lambda v, i: None if v is None else v[i:i + per]
        pieces = [_render_rays_fused(ray_batch[i:i + per], network_fn, N_samples, perturb, N_importance, network_fine, raw_noise_std,
                                     pytest, test_time, args, cut(t_rand, i), cut(u, i), cut(noise, i), return_aux)
                  for i in range(0, n_all, per)]
        out = {k: torch.cat([p_[k] for p_ in pieces], 0) for k in pieces[0] if k != "_aux"}
        if "_aux" in pieces[0]:
            out["_aux"] = {k: torch.cat([p_["_aux"][k] for p_ in pieces], 0) for k in pieces[0]["_aux"]}
        return out
    N_rays, width = ray_batch.shape
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, 8:11] if width > 8 else None
    img_idxs = ray_batch[:, 11:]
    dev = ray_batch.device

    if perturb > 0. and t_rand is None:
        t_rand = torch.rand(N_rays, N_samples, device=dev)                        # rendering.py:110
    if lindisp:                                                                   # rendering.py:100, :104-112: glue, as written
        near, far = ray_batch.detach()[:, 6:7], ray_batch.detach()[:, 7:8]
        t_vals = ops.linspace01(N_samples, dev)
        z_vals = (1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)).expand(N_rays, N_samples)
        if perturb > 0.:
            mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper, lower = torch.cat([mids, z_vals[..., -1:]], -1), torch.cat([z_vals[..., :1], mids], -1)
            z_vals = lower + (upper - lower) * t_rand
        z_vals = z_vals.contiguous()
    else:
        z_vals = ops.sample_coarse(ray_batch.detach()[:, 6], ray_batch.detach()[:, 7], width, N_rays, N_samples,
                                   t_rand if perturb > 0. else None)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]      # rendering.py:114

    store_rgb = N_importance == 0
    with ops.tiled_raw():
        raw = network_query_fn(pts, viewdirs, None, network_fn, 'coarse', False, test_time=test_time, store_rgb=store_rgb)
    if isinstance(raw, ops.TiledRaw):
        raw.private = raw.single                     # consumed by the compositing below and by nothing else
    if noise is None and raw_noise_std > 0. and not (test_time and not store_rgb):
        noise = torch.randn(N_rays, N_samples, device=dev) * raw_noise_std        # nerfh_nff.py:67
    rgb_map, feat_map, disp_map, acc_map, weights, depth_map, _, _ = raw2outputs_NeRFH_NFF(
        raw, z_vals, raw_noise_std=raw_noise_std, white_bkgd=white_bkgd, test_time=test_time, typ="coarse",
        store_rgb=store_rgb, noise=noise)

    aux = {}
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0, feat_map_0 = rgb_map, disp_map, acc_map, feat_map
        det = (perturb == 0.)
        if not det and u is None:
            u = torch.rand(N_rays, N_importance, device=dev)                      # rendering.py:36
        if pytest:
            import numpy as np
            np.random.seed(0)
            u = None if det else torch.Tensor(np.random.rand(N_rays, N_importance)).to(dev)
        z_fine, z_samples, inds = ops.sample_fine(z_vals, weights.detach(), N_importance, None if det else u)
        if args.use_fine_only:
            z_vals_f = z_samples
        else:
            z_vals_f = z_fine
        if return_aux:
            aux = dict(z_coarse=z_vals, z_samples=z_samples, inds=inds, z_fine=z_vals_f, weights_coarse=weights)
        z_vals = z_vals_f
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        output_transient = bool(args.NeRFW)
        with ops.tiled_raw():
            raw = network_query_fn(pts, viewdirs, img_idxs, network_fine, 'fine', output_transient,
                                   test_time=test_time, store_rgb=store_rgb)
        if isinstance(raw, ops.TiledRaw):
            raw.private = raw.single
        rgb_map, feat_map, disp_map, acc_map, weights, depth_map, transient_sigmas, beta = raw2outputs_NeRFH_NFF(
            raw, z_vals, raw_noise_std=raw_noise_std, output_transient=output_transient,
            beta_min=network_fine.beta_min, white_bkgd=white_bkgd, test_time=test_time, typ="fine",
            transient_at_test=args.transient_at_test)

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
    if args.nerfh_nff:
        ret['feat_map'] = feat_map
    if (N_importance > 0 and test_time) or (N_importance == 0):
        pass
    elif N_importance > 0:
        ret['rgb0'], ret['disp0'], ret['acc0'] = rgb_map_0, disp_map_0, acc_map_0
        ret['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)
        if args.NeRFW:
            ret['transient_sigmas'] = transient_sigmas
            ret['beta'] = beta
        if args.nerfh_nff and feat_map_0 is not None:
            ret['feat0'] = feat_map_0
    if return_aux:
        aux.update(depth_map=depth_map, weights_fine=weights)
        for k, v in aux.items():
            ret['aux_' + k] = v
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Drop-in for rendering.py:182-195.  Per-ray extras (t_rand, u, noise) are chunked with the rays."""
    n = rays_flat.shape[0]
    per_ray = {k: kwargs.pop(k) for k in ("t_rand", "u", "noise") if k in kwargs}
    if n <= chunk:
        return render_rays(rays_flat, **per_ray, **kwargs)
    all_ret = {}
    for i in range(0, n, chunk):
        sub = {k: (None if v is None else v[i:i + chunk]) for k, v in per_ray.items()}
        ret = render_rays(rays_flat[i:i + chunk], **sub, **kwargs)
        for k, v in ret.items():
            all_ret.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in all_ret.items()}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, img_idx=torch.Tensor(0), **kwargs):
    """Drop-in for rendering.py:197-243: returns [rgb_map, disp_map, acc_map, extras]."""
    if ndc:
        raise RuntimeError("nefes_b200: ndc=True is not on the NeFeS path (render_kwargs set ndc=False)")
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w)
    else:
        rays_o, rays_d = rays
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, near, far], -1)
    if use_viewdirs:
        rays = torch.cat([rays, viewdirs], -1)
    img_idx = img_idx.to(rays.device)
    if img_idx.shape[0] != rays.shape[0]:
        img_idx = img_idx.repeat(rays.shape[0], 1)
    rays = torch.cat([rays, img_idx.to(rays.dtype)], 1)
    all_ret = batchify_rays(rays, chunk, **kwargs)
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    return [all_ret[k] for k in k_extract] + [{k: v for k, v in all_ret.items() if k not in k_extract}]
