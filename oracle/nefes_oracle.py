"""CPU oracle for the NeFeS render hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the algorithm the
reference implements in script/models/{ray_utils,rendering,nerfh_nff}.py.  It is the
checker the CUDA engine is compared with; it is never the thing measured or shipped.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it.  The product package (nefes_b200/) never imports oracle/.

Parity status: PINNED.  oracle/make_golden.py imports the unmodified reference from
/root/reference (in the build container), runs both on identical seeded inputs,
asserts bit-equality on CPU fp32 and writes the vectors under tests/golden/;
tests/test_oracle_golden.py re-checks the oracle against those vectors everywhere.

Every function cites the reference lines it follows (paths relative to
/root/reference/).  Random numbers are never drawn here: callers pass `t_rand`,
`noise` and `u` explicitly (see draw_train_randoms for the reference's draw order).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

FEAT_CH = 128          # script/models/nerfh_nff.py:21
XYZ_FREQS = 10         # script/models/options.py:99
DIR_FREQS = 4          # script/models/options.py:100
HIDDEN = 128           # script/models/options.py:31
DEPTH = 8              # script/models/options.py:30
SKIP_AT = 4            # script/models/nerfh_nff.py:640
LAST_DELTA = 1e2       # script/models/nerfh_nff.py:56


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def init_field(typ: str, width: int = HIDDEN, depth: int = DEPTH,
               xyz_ch: int = 63, dir_ch: int = 27, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init parameters with the reference constructor's RNG order and key names.

    script/models/nerfh_nff.py:446 (manual_seed(0)), :469-505 (layer creation order).
    Returns a flat dict keyed exactly like NeRFH_NFF.state_dict() for the MLP part
    (SURVEY.md §8a row a14).
    """
    torch.manual_seed(0)
    out: Dict[str, torch.Tensor] = {}

    def lin(name: str, fan_in: int, fan_out: int):
        layer = torch.nn.Linear(fan_in, fan_out)
        out[name + ".weight"] = layer.weight.detach().to(dtype).clone()
        out[name + ".bias"] = layer.bias.detach().to(dtype).clone()

    for i in range(depth):
        fan_in = xyz_ch if i == 0 else (width + xyz_ch if i == SKIP_AT else width)
        lin(f"xyz_encoding_{i + 1}.0", fan_in, width)
    lin("xyz_encoding_final", width, width)
    lin("dir_encoding.0", width + dir_ch, width // 2)
    lin("static_sigma.0", width, 1)
    lin("static_rgb.0", width // 2, 3 + FEAT_CH)
    if typ == "fine":
        lin("transient_encoding.0", width + dir_ch, width // 2)
        lin("transient_encoding.2", width // 2, width // 2)
        lin("transient_encoding.4", width // 2, width // 2)
        lin("transient_sigma.0", width // 2, 1)
        lin("transient_rgb.0", width // 2, 3)
        lin("transient_beta.0", width // 2, 1)
    return out


def clone_params(p: Dict[str, torch.Tensor], dtype=None, requires_grad=False):
    q = {}
    for k, v in p.items():
        t = v.detach().clone()
        if dtype is not None:
            t = t.to(dtype)
        q[k] = t.requires_grad_(requires_grad)
    return q


# --------------------------------------------------------------------------------------
# rays
# --------------------------------------------------------------------------------------
def camera_rays(H: int, W: int, focal: float, c2w: torch.Tensor):
    """script/models/ray_utils.py:5-16.  No half-pixel offset; principal point (W/2, H/2)."""
    cols = torch.linspace(0, W - 1, W, dtype=c2w.dtype)
    rows = torch.linspace(0, H - 1, H, dtype=c2w.dtype)
    px = cols[None, :].expand(H, W)
    py = rows[:, None].expand(H, W)
    cam = torch.stack([(px - W * .5) / focal, -(py - H * .5) / focal, -torch.ones_like(px)], -1)
    rays_d = torch.sum(cam[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def camera_rays_batch(H: int, W: int, focal: float, c2w: torch.Tensor):
    """script/models/ray_utils.py:46-59."""
    o, d = zip(*[camera_rays(H, W, focal, c2w[b]) for b in range(c2w.shape[0])])
    return torch.stack(o), torch.stack(d)


def pack_rays(rays_o, rays_d, near: float, far: float, hist: torch.Tensor):
    """script/models/rendering.py:209-235 -> [N, 3+3+1+1+3+10]."""
    rays_o = rays_o.reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    view = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    ones = torch.ones_like(rays_d[..., :1])
    if hist.shape[0] != rays_o.shape[0]:
        hist = hist.repeat(rays_o.shape[0], 1)
    return torch.cat([rays_o, rays_d, near * ones, far * ones, view, hist.to(rays_o.dtype)], -1)


# --------------------------------------------------------------------------------------
# sampling
# --------------------------------------------------------------------------------------
def coarse_depths(near, far, n: int, t_rand: Optional[torch.Tensor], dtype=torch.float32, lindisp=False):
    """script/models/rendering.py:96-112.  near/far: [N,1].  lindisp: linear in disparity (:100)."""
    t = torch.linspace(0., 1., steps=n).to(dtype)
    if lindisp:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    else:
        z = near * (1. - t) + far * t
    z = z.expand(near.shape[0], n)
    if t_rand is not None:                                     # perturb > 0
        mid = .5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], -1)
        lo = torch.cat([z[..., :1], mid], -1)
        z = lo + (hi - lo) * t_rand
    return z


def pdf_to_cdf(weights):
    """script/models/rendering.py:26-29."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def invert_cdf(bins, cdf, u):
    """script/models/rendering.py:49-64 -> (samples, inds).  inds is the raw
    searchsorted(right=True) result in [0, len(cdf)] (the 'sample indices')."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf.detach().contiguous(), u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    denom = c_hi - c_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - c_lo) / denom
    return b_lo + t * (b_hi - b_lo), inds


def importance_depths(bins, weights, n: int, u: Optional[torch.Tensor]):
    """script/models/rendering.py:23-66.  u=None -> det=True (linspace)."""
    cdf = pdf_to_cdf(weights)
    if u is None:
        u = torch.linspace(0., 1., steps=n).to(cdf.dtype).expand(list(cdf.shape[:-1]) + [n])
    samples, inds = invert_cdf(bins, cdf, u)
    return samples, inds, cdf


# --------------------------------------------------------------------------------------
# field (positional encoding + MLP)
# --------------------------------------------------------------------------------------
def freq_encode(x, n_freqs: int):
    """script/models/nerfh_nff.py:241-270, log-sampled bands 2^0..2^(L-1), [x, sin, cos, ...]."""
    parts = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        parts += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(parts, -1)


def field_forward(P: Dict[str, torch.Tensor], enc_xyz, enc_dir=None, mode: str = "full", q=None):
    """script/models/nerfh_nff.py:525-576.

    mode 'sigma'  -> [M,1]   (sigma_only=True)
    mode 'static' -> [M,132] (output_transient=False)
    mode 'full'   -> [M,137] (fine net with transient heads)

    q (optional, not in the reference): rounding applied to every tensor that is a matmul operand
    (inputs and hidden activations).  q = bf16 round-trip emulates the engine's bf16 tensor-core path
    (bf16 operands, fp32 accumulate) so that path can be checked against like-for-like arithmetic.
    """
    q = q or (lambda t: t)
    enc_xyz = q(enc_xyz)
    h = enc_xyz
    for i in range(DEPTH):
        if i == SKIP_AT:
            h = torch.cat([enc_xyz, h], 1)
        h = q(F.relu(F.linear(h, P[f"xyz_encoding_{i + 1}.0.weight"], P[f"xyz_encoding_{i + 1}.0.bias"])))
    sigma = F.softplus(F.linear(h, P["static_sigma.0.weight"], P["static_sigma.0.bias"]))
    if mode == "sigma":
        return sigma
    fin = q(F.linear(h, P["xyz_encoding_final.weight"], P["xyz_encoding_final.bias"]))
    both = torch.cat([fin, q(enc_dir)], 1)
    dh = q(F.relu(F.linear(both, P["dir_encoding.0.weight"], P["dir_encoding.0.bias"])))
    rgbf = F.linear(dh, P["static_rgb.0.weight"], P["static_rgb.0.bias"])      # 131 ch, no activation
    static = torch.cat([rgbf, sigma], 1)
    if mode == "static":
        return static
    t = both
    for j in (0, 2, 4):
        t = q(F.relu(F.linear(t, P[f"transient_encoding.{j}.weight"], P[f"transient_encoding.{j}.bias"])))
    t_sigma = F.softplus(F.linear(t, P["transient_sigma.0.weight"], P["transient_sigma.0.bias"]))
    t_rgb = torch.sigmoid(F.linear(t, P["transient_rgb.0.weight"], P["transient_rgb.0.bias"]))
    t_beta = F.softplus(F.linear(t, P["transient_beta.0.weight"], P["transient_beta.0.bias"]))
    return torch.cat([static, t_rgb, t_sigma, t_beta], 1)


def query_field(P, pts, viewdirs, typ: str, output_transient: bool, test_time: bool, q=None):
    """script/models/nerfh_nff.py:168-231 (netchunk loop dropped: one chunk)."""
    n, s = pts.shape[:2]
    flat = pts.reshape(-1, 3)
    ex = freq_encode(flat, XYZ_FREQS)
    if typ == "coarse" and test_time:
        return field_forward(P, ex, mode="sigma", q=q).reshape(n, s, -1)
    ed = freq_encode(viewdirs[:, None].expand(pts.shape).reshape(-1, 3), DIR_FREQS)
    mode = "full" if (typ == "fine" and output_transient) else "static"
    return field_forward(P, ex, ed, mode, q=q).reshape(n, s, -1)


def bf16_round(t):
    """Round-to-nearest-even to bf16 and back (differentiable as identity)."""
    return t + (t.detach().bfloat16().float() - t.detach())


def bf16_weights(P):
    """Weights rounded to bf16 (what the tensor-core path multiplies with); biases stay fp32.  The
    returned leaves are the ORIGINAL fp32 tensors' rounded copies with straight-through gradients."""
    return {k: (bf16_round(v) if k.endswith(".weight") else v) for k, v in P.items()}


# --------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------
@dataclass
class Composite:
    rgb: Optional[torch.Tensor]
    feat: Optional[torch.Tensor]
    disp: Optional[torch.Tensor]
    acc: torch.Tensor
    weights: torch.Tensor
    depth: Optional[torch.Tensor]
    transient_sigmas: Optional[torch.Tensor]
    beta: Optional[torch.Tensor]

    def astuple(self):
        return (self.rgb, self.feat, self.disp, self.acc, self.weights, self.depth,
                self.transient_sigmas, self.beta)


def _exclusive_cumprod(one_minus_alpha):
    lead = torch.ones_like(one_minus_alpha[:, :1])
    return torch.cumprod(torch.cat([lead, one_minus_alpha], -1)[:, :-1], -1)


def composite(raw, z, noise=None, output_transient=False, beta_min=0.1, test_time=False,
              typ="coarse", store_rgb=False, transient_at_test=False, white_bkgd=False) -> Composite:
    """script/models/nerfh_nff.py:25-166.  `noise` = randn*raw_noise_std (or None = 0).  white_bkgd acts in the one branch
    where the reference still applies it (:126-127, transient compositing of the fine pass); elsewhere it is commented out."""
    sigma_only = typ == "coarse" and test_time and not store_rgb
    if sigma_only:
        s_sig, t_sig = raw[..., 0], None
    else:
        c = raw.shape[-1] - (6 if output_transient else 1)
        s_rgb, s_sig = raw[..., :c], raw[..., c]
        if output_transient:
            t_rgb, t_sig, t_beta = raw[..., c + 1:c + 4], raw[..., c + 4], raw[..., c + 5]
        else:
            t_sig = None
    delta = torch.cat([z[:, 1:] - z[:, :-1], LAST_DELTA * torch.ones_like(z[:, :1])], -1)   # :55-60
    if output_transient:
        a_s = 1 - torch.exp(-delta * s_sig)
        a_t = 1 - torch.exp(-delta * t_sig)
        a = 1 - torch.exp(-delta * (s_sig + t_sig))
    else:
        dens = s_sig if noise is None else s_sig + noise
        a = 1 - torch.exp(-delta * dens)
    T = _exclusive_cumprod(1 - a)                                                          # :70-71
    w = a * T
    acc = w.sum(-1)
    if sigma_only:                                                                           # :83-89
        return Composite(None, None, None, acc, w, None, t_sig, None)
    if output_transient:
        if test_time and not transient_at_test:                                              # :92-117
            w_s = a_s * _exclusive_cumprod(1 - a_s)
            rgb = (w_s[..., None] * s_rgb[..., :3]).sum(1)
            feat = (w_s.detach()[..., None] * s_rgb[..., 3:]).sum(1)
            depth = (w_s * z).sum(-1)
            disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / w_s.sum(-1))
            return Composite(rgb, feat, disp, acc, w_s, depth, t_sig, torch.zeros_like(acc))
        w_s, w_t = a_s * T, a_t * T
        rgb_s = (w_s[..., None] * s_rgb[..., :3]).sum(1)
        if white_bkgd:
            rgb_s = rgb_s + (1 - acc[..., None])                                             # :126-127
        rgb = rgb_s + (w_t[..., None] * t_rgb).sum(1)                                        # :119-150
        feat = (w_s.detach()[..., None] * s_rgb[..., 3:]).sum(1)                            # :122-125
        beta = (w_t * t_beta).sum(-1) + beta_min                                             # :133-137
    else:
        rgb = (w[..., None] * s_rgb[..., :3]).sum(1)                                         # :153-157
        feat = (w.detach()[..., None] * s_rgb[..., 3:]).sum(1)
        beta = torch.zeros_like(acc)
    depth = (w * z).sum(-1)                                                                  # :164-165
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / w.sum(-1))
    return Composite(rgb, feat, disp, acc, w, depth, t_sig, beta)


# --------------------------------------------------------------------------------------
# render_rays / render
# --------------------------------------------------------------------------------------
def march_rays(ray_batch, P_coarse, P_fine, n_coarse=64, n_fine=64, test_time=False,
               t_rand=None, noise=None, u=None, transient=True, transient_at_test=True,
               beta_min=0.1, return_aux=False, emulate_bf16=False, lindisp=False, white_bkgd=False):
    """script/models/rendering.py:68-180 with args.nerfh_nff=True, use_fine_only=False,
    NeRFW=`transient`.  Train mode needs t_rand [N,n_coarse] and u [N,n_fine];
    test mode (perturb=0) uses neither.
    emulate_bf16 (not in the reference): both fields run with the rounding points of the engine's tensor-core path
    (weights and every matmul operand rounded to bf16, fp32 accumulation; see field_forward's `q`)."""
    q = None
    if emulate_bf16:
        P_coarse, P_fine, q = bf16_weights(P_coarse), bf16_weights(P_fine), bf16_round
    o, d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    view = ray_batch[:, 8:11]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z_c = coarse_depths(near, far, n_coarse, None if test_time else t_rand, ray_batch.dtype, lindisp=lindisp)
    pts = o[:, None, :] + d[:, None, :] * z_c[..., None]
    raw_c = query_field(P_coarse, pts, view, "coarse", False, test_time, q=q)
    c0 = composite(raw_c, z_c, noise=noise, test_time=test_time, typ="coarse", white_bkgd=white_bkgd)
    mids = .5 * (z_c[..., 1:] + z_c[..., :-1])
    z_s, inds, cdf = importance_depths(mids, c0.weights[..., 1:-1], n_fine, None if test_time else u)
    z_s = z_s.detach()
    z_f, _ = torch.sort(torch.cat([z_c, z_s], -1), -1)
    pts_f = o[:, None, :] + d[:, None, :] * z_f[..., None]
    raw_f = query_field(P_fine, pts_f, view, "fine", transient, test_time, q=q)
    c1 = composite(raw_f, z_f, output_transient=transient, beta_min=beta_min, test_time=test_time,
                   typ="fine", transient_at_test=transient_at_test, white_bkgd=white_bkgd)
    ret = {"rgb_map": c1.rgb, "disp_map": c1.disp, "acc_map": c1.acc, "feat_map": c1.feat}
    if not test_time:                                                                        # rendering.py:163-176
        ret.update(rgb0=c0.rgb, disp0=c0.disp, acc0=c0.acc,
                   z_std=torch.std(z_s, dim=-1, unbiased=False))
        if transient:
            ret.update(transient_sigmas=c1.transient_sigmas, beta=c1.beta)
        ret["feat0"] = c0.feat
    if return_aux:
        ret["_aux"] = dict(z_coarse=z_c, z_samples=z_s, inds=inds, cdf=cdf, z_fine=z_f,
                           weights_coarse=c0.weights, weights_fine=c1.weights,
                           depth_map=c1.depth, raw_coarse=raw_c, raw_fine=raw_f)
    return ret


def render(H, W, focal, P_coarse, P_fine, *, rays=None, c2w=None, near=0., far=1.,
           hist=None, chunk=32768, **kw):
    """script/models/rendering.py:197-243 with use_viewdirs=True, ndc=False."""
    if c2w is not None:
        rays_o, rays_d = camera_rays(H, W, focal, c2w)
    else:
        rays_o, rays_d = rays
    if hist is None:
        hist = torch.zeros(1, 10)
    batch = pack_rays(rays_o, rays_d, near, far, hist)
    pieces = []
    for i in range(0, batch.shape[0], chunk):
        sl = slice(i, i + chunk)
        sub = {k: (v[sl] if isinstance(v, torch.Tensor) and v.shape[:1] == batch.shape[:1] else v)
               for k, v in kw.items()}
        pieces.append(march_rays(batch[sl], P_coarse, P_fine, **sub))
    out = {k: torch.cat([p[k] for p in pieces], 0) for k in pieces[0] if k != "_aux"}
    if "_aux" in pieces[0]:
        out["_aux"] = {k: torch.cat([p["_aux"][k] for p in pieces], 0) for k in pieces[0]["_aux"]}
    return out


def draw_train_randoms(n_rays: int, n_coarse=64, n_fine=64, seed: int = 0):
    """Reproduce the reference's CPU RNG consumption order for one render_rays call in
    train mode: rand[N,64] (rendering.py:110), randn_like[N,64] (nerfh_nff.py:67, drawn even
    when raw_noise_std == 0), rand[N,64] (rendering.py:36)."""
    torch.manual_seed(seed)
    t_rand = torch.rand(n_rays, n_coarse)
    _noise = torch.randn(n_rays, n_coarse)
    u = torch.rand(n_rays, n_fine)
    return t_rand, u


# --------------------------------------------------------------------------------------
# callers that the benchmarks need (loss, pose chain) -- restated, small
# --------------------------------------------------------------------------------------
def nerfw_loss(ret, target_rgb, lambda_u=0.01):
    """script/models/losses.py:112-132 (coef=1)."""
    c_l = 0.5 * ((ret["rgb0"] - target_rgb) ** 2).mean()
    f_l = ((ret["rgb_map"] - target_rgb) ** 2 / (2 * ret["beta"].unsqueeze(1) ** 2)).mean()
    b_l = 3 + torch.log(ret["beta"]).mean()
    s_l = lambda_u * ret["transient_sigmas"].mean()
    return c_l + f_l + b_l + s_l


def color_feature_fusion_nerfw_loss(results, targets, switch_on=True, color_only_switch=False, L1_loss=True, lambda_u=0.01):
    """script/models/losses.py:134-173 (ColorFeatureFusionNerfWLoss, coef=1).  `results` uses the caller's key names
    (run_nefes.py:217-231): rgb_fine, rgb_coarse, beta, transient_sigmas, feat_fine, [feat_coarse], [feat_fusion]."""
    loss = nerfw_loss({"rgb0": results["rgb_coarse"], "rgb_map": results["rgb_fine"], "beta": results["beta"],
                       "transient_sigmas": results["transient_sigmas"]}, targets["rgb"], lambda_u)
    if color_only_switch:
        return loss
    f = (lambda a, b: (a - b).abs().mean()) if L1_loss else (lambda a, b: ((a - b) ** 2).mean())
    loss_f = f(results["feat_fine"], targets["feat"])
    if "feat_coarse" in results:
        loss_f = loss_f + f(results["feat_coarse"], targets["feat"])
    if switch_on:
        return loss, loss_f, f(results["feat_fusion"], targets["feat"])
    return loss, loss_f


IMG_MEAN, IMG_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)                # script/models/nerfh_nff.py:359-360


def fusion_net(P, rgb, feat, B, H, W, training=True, no_bn=False, residual=False, running=None, momentum=0.1, eps=1e-5):
    """script/models/nerfh_nff.py:578-603 (run_fusion_net) + :356-418 (FusionNet.forward): rgb [B*H*W,3], feat [B*H*W,128] ->
    feature_output [B,128,H,W].  P: the FusionNet state_dict (net.0/2/4/6 convolutions, net.7 BatchNorm2d); `running` =
    (running_mean, running_var) tensors, updated in place when training (torch semantics)."""
    x = torch.cat([rgb.reshape(B, H, W, 3).permute(0, 3, 1, 2), feat.reshape(B, H, W, -1).permute(0, 3, 1, 2)], 1)
    mean, std = x.new_tensor(IMG_MEAN), x.new_tensor(IMG_STD)
    x = torch.cat([(x[:, :3] - mean[:, None, None]) / std[:, None, None], x[:, 3:]], 1)
    h = F.relu(F.conv2d(x, P["net.0.weight"], P["net.0.bias"], padding=1))
    h = F.relu(F.conv2d(h, P["net.2.weight"], P["net.2.bias"], padding=1))
    h = F.relu(F.conv2d(h, P["net.4.weight"], P["net.4.bias"], padding=1))
    y = F.conv2d(h, P["net.6.weight"], P["net.6.bias"], padding=2)
    if not no_bn:
        rm, rv = running if running is not None else (P["net.7.running_mean"], P["net.7.running_var"])
        y = F.batch_norm(y, rm, rv, P["net.7.weight"], P["net.7.bias"], training, momentum, eps)
    return x[:, 3:] + y if residual else y


def exposure_mlp(params, hist):
    """The exposure network (script/models/nerfh_nff.py:511-522): tiny-cuda-nn FullyFusedMLP 10 -> 32 -> 32 -> 32 -> 12, ReLU, no
    biases.  tiny-cuda-nn is not vendored and not pinned by the reference: PARITY UNPINNED -- restated from its published
    layout (flat buffer of [out, in] row-major matrices, 10 inputs padded to 16 with ONES, 12 outputs padded to 16), fp32."""
    h = F.pad(hist.long().to(params.dtype), (0, 6), value=1.0)
    off = 0
    for li, (o, i) in enumerate(((32, 16), (32, 32), (32, 32), (16, 32))):
        h = h @ params[off:off + o * i].view(o, i).t()
        off += o * i
        if li < 3:
            h = torch.relu(h)
    return h[:, :12]


def affine_color(params, rgb, hist, B):
    """script/models/nerfh_nff.py:605-626: rgb [B*N,3] -> sigmoid(K_b rgb + bias_b), [K_b | bias_b] = exposure_mlp(hist_b)."""
    a = exposure_mlp(params, hist)
    K, bias = a[:, :9].reshape(-1, 3, 3), a[:, 9:].reshape(-1, 3, 1)
    out = torch.bmm(K, rgb.reshape(B, -1, 3).transpose(1, 2)) + bias
    return torch.sigmoid(out.transpose(1, 2).reshape(-1, 3))


def cosine_feature_loss(feat_render, feat_target):
    """script/dm/DFM_pose_refine.py:236-255, per_pixel=False, inputs [C, N]."""
    return 1 - F.cosine_similarity(feat_render, feat_target, dim=1, eps=1e-6).mean()


def so3_exp(r):
    """script/utils/lie_group_helper.py:60-70 (axis-angle -> rotation, Rodrigues)."""
    zero = torch.zeros((), dtype=r.dtype)
    K = torch.stack([torch.stack([zero, -r[2], r[1]]),
                     torch.stack([r[2], zero, -r[0]]),
                     torch.stack([-r[1], r[0], zero])])
    n = r.norm() + 1e-15
    return torch.eye(3, dtype=r.dtype) + (torch.sin(n) / n) * K + ((1 - torch.cos(n)) / n ** 2) * (K @ K)


def learn_pose_c2w(r, t, init_c2w):
    """script/models/poses.py:25-50 with lietorch=False (make_c2w path): R = Exp(r) @ R0,
    trans = t + t0.  Returns [3,4]."""
    R = so3_exp(r) @ init_c2w[:3, :3]
    tr = t + init_c2w[:3, 3]
    return torch.cat([R, tr[:, None]], 1)


def pose_error(c2w_a, c2w_b):
    """script/dm/pose_model.py:75-92 style: (metres, degrees)."""
    dt = float(torch.linalg.norm(c2w_a[:3, 3] - c2w_b[:3, 3]))
    Rrel = c2w_a[:3, :3].double() @ c2w_b[:3, :3].double().T
    cosang = max(-1.0, min(1.0, (float(torch.trace(Rrel)) - 1) / 2))
    return dt, math.degrees(math.acos(cosang))
