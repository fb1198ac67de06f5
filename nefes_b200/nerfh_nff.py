"""The reference's script/models/nerfh_nff.py interface for the render hot path on the B200 kernels: same names,
argument meaning and return conventions (the drop-in boundary -- `create_nerf`'s dict keys and the module's attribute /
state_dict names ARE the interface and are restated as such); the arithmetic behind them is the engine's.

    raw2outputs_NeRFH_NFF   nerfh_nff.py:25-166
    run_network_NeRFH_NFF   nerfh_nff.py:168-231
    get_embedder / Embedder nerfh_nff.py:234-354
    NeRFH_NFF               nerfh_nff.py:421-626
    create_nerf             nerfh_nff.py:628-737
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops

FEATURE_DIM = 128
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)

# default arithmetic of the MLP: "fp32" (exact SIMT path), "tf32" (tcgen05 kind::tf32 GEMMs under the fp32 path's structure:
# the <= 1e-3 tensor-core path) or "bf16" (fused tcgen05 layer chains, what bench.py times); see NeRFH_NFF.precision
DEFAULT_PRECISION = os.environ.get("NEFES_PRECISION", "fp32")
_PREC = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "tf32": L.PREC_TF32}


# ------------------------------------------------------------------------------------------------
def raw2outputs_NeRFH_NFF(raw, z_vals, raw_noise_std=0, output_transient=False, beta_min=0.1, white_bkgd=False,
                          test_time=False, typ="coarse", store_rgb=False, transient_at_test=False, noise=None):
    """Drop-in for nerfh_nff.py:25.  Returns (rgb_map, features_map, disp_map, acc_map, weights,
    depth_map, transient_sigmas, beta).  `noise` (extra, optional) supplies randn*raw_noise_std
    explicitly for RNG parity; otherwise it is drawn on the device when raw_noise_std > 0.
    white_bkgd acts where the reference still applies it (:126-127: the transient compositing adds 1 - acc to the static
    colour); in the other branches the reference has it commented out (:104, :158) and so it has no effect there."""
    if typ == "coarse" and test_time and not store_rgb:                      # :33-35, :83-89
        acc, weights = ops.composite(raw[..., :1], z_vals, None, L.COMP_SIGMA, beta_min)
        return None, None, None, acc, weights, None, None, None
    if not isinstance(raw, ops.TiledRaw) and raw.shape[-1] in (4, 9) and raw.shape[-1] == (9 if output_transient else 4):
        # front-end B (nerfh_tcnn.py): rgb + sigma [+ transient rgb, sigma, beta] and NO feature channels (ch_rgbs = 3, :38-44).
        # The compositing kernels are built for the 3 + 128-channel head: the raw goes through them with the 128 feature
        # channels zero-filled, and the (empty) feature map comes back with zero channels, as the reference's slicing gives.
        pad = raw.new_zeros(raw.shape[:2] + (137 if output_transient else 132,))
        pad[..., :3] = raw[..., :3]
        pad[..., 131:131 + raw.shape[-1] - 3] = raw[..., 3:]
        out = list(raw2outputs_NeRFH_NFF(pad, z_vals, raw_noise_std, output_transient, beta_min, white_bkgd, test_time, typ,
                                         store_rgb, transient_at_test, noise))
        out[1] = out[1][..., :0]
        return tuple(out)
    if output_transient:
        if raw.shape[-1] != 137:
            raise RuntimeError(f"nefes_b200: transient compositing expects 137 channels, got {raw.shape[-1]}")
        mode = L.COMP_TRANSIENT_STATIC_ONLY if (test_time and not transient_at_test) else L.COMP_TRANSIENT
        rgb, feat, disp, acc, weights, depth, beta, tsig = ops.composite(raw, z_vals, None, mode, beta_min)
        if white_bkgd and mode == L.COMP_TRANSIENT:
            rgb = rgb + (1 - acc[..., None])
        return rgb, feat, disp, acc, weights, depth, tsig, beta
    if raw.shape[-1] != 132:
        raise RuntimeError(f"nefes_b200: static compositing expects 132 channels, got {raw.shape[-1]}")
    if noise is None and raw_noise_std > 0.:
        noise = torch.randn(raw.shape[:2], device=raw.device) * raw_noise_std
    rgb, feat, disp, acc, weights, depth, beta = ops.composite(raw, z_vals, noise, L.COMP_STATIC, beta_min)
    return rgb, feat, disp, acc, weights, depth, None, beta


def run_network_NeRFH_NFF(inputs, viewdirs, ts, fn, embed_fn=None, embeddirs_fn=None, typ="coarse",
                          output_transient=False, netchunk=1024 * 64, test_time=False, store_rgb=False):
    """Drop-in for nerfh_nff.py:168.  `ts` is accepted and ignored, as in the reference.  The
    positional encodings are fused into the field kernel, so embed_fn / embeddirs_fn are unused;
    netchunk bounds the rays handed to one launch (activation workspace), not the arithmetic."""
    n_rays, n_samples = inputs.shape[0], inputs.shape[1]
    if typ == "coarse" and test_time:
        mode = L.MODE_SIGMA
    elif typ == "fine" and output_transient:
        mode = L.MODE_FULL
    else:
        mode = L.MODE_STATIC
    rays_per_chunk = max(1, int(netchunk) // n_samples)
    if n_rays <= rays_per_chunk:
        return fn.query(inputs, viewdirs, mode)
    if (rays_per_chunk * n_samples) % 128:                    # keep chunk boundaries on tile boundaries
        rays_per_chunk = max(1, rays_per_chunk - rays_per_chunk % 128)
    outs = [fn.query(inputs[i:i + rays_per_chunk], None if viewdirs is None else viewdirs[i:i + rays_per_chunk], mode)
            for i in range(0, n_rays, rays_per_chunk)]
    if isinstance(outs[0], ops.TiledRaw):
        return ops.TiledRaw.cat(outs)
    return torch.cat(outs, 0)


class StandardQuery:
    """The `network_query_fn` closure create_nerf builds (nerfh_nff.py:667-675) as an object: same call, and render_rays can
    recognise it (`nefes_standard`) and hand the whole path to nefes_render_rays_fwd/_bwd as ONE engine call whenever the
    rays fit one netchunk.  A hand-written query function keeps the staged path."""
    nefes_standard = True

    def __init__(self, netchunk=1024 * 64, embed_fn=None, embeddirs_fn=None):
        self.netchunk, self.embed_fn, self.embeddirs_fn = int(netchunk), embed_fn, embeddirs_fn

    def __call__(self, inputs, viewdirs, ts, network_fn, typ, output_transient, test_time, store_rgb):
        return run_network_NeRFH_NFF(inputs, viewdirs, ts, network_fn, embed_fn=self.embed_fn, embeddirs_fn=self.embeddirs_fn,
                                     typ=typ, output_transient=output_transient, netchunk=self.netchunk,
                                     test_time=test_time, store_rgb=store_rgb)


# ------------------------------------------------------------------------------------------------
class Embedder:
    """nerfh_nff.py:234-270: [x, sin(2^k x), cos(2^k x)]_k with log-sampled bands."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.N_freqs = kwargs["num_freqs"]
        d = kwargs["input_dims"]
        if d != 3 or not kwargs.get("include_input", True) or not kwargs.get("log_sampling", True):
            raise RuntimeError("nefes_b200: only the 3-D, include_input, log-sampled embedder is built")
        self.out_dim = d + 2 * d * self.N_freqs

    def embed(self, inputs):
        if self.kwargs["max_freq_log2"] == 0:
            return inputs
        return ops.encode_pe(inputs, self.N_freqs)


def get_embedder(multires, i=0, reduce_mode=-1, epochToMaxFreq=-1):
    """nerfh_nff.py:303-354 (reduce_mode -1: the paper default, the only one the configs use)."""
    if i == -1:
        return nn.Identity(), 3
    if reduce_mode not in (-1,):
        raise RuntimeError("nefes_b200: reduce_embedding modes 0/1/2 are not built")
    obj = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                   log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return (lambda x, eo=obj: eo.embed(x)), obj.out_dim, obj


# ------------------------------------------------------------------------------------------------
class FusionNet(nn.Module):
    """nerfh_nff.py:356-418 -- the CNN right after the render path.  The module owns the parameters under the
    reference's keys (net.0 ... net.7) so checkpoints load; NeRFH_NFF.run_fusion_net runs it on the engine (fusion.py ->
    nefes_fusion_fwd/_bwd); this torch forward remains for NCHW callers and as a CPU mirror."""
    mean = [0.485, 0.456, 0.406]
    std = [0.229, 0.224, 0.225]

    def __init__(self, f_dim, fusion_residule=False, no_BN=False):
        super().__init__()
        self.fusion_residule, self.no_BN = fusion_residule, no_BN
        layers = [nn.Conv2d(3 + f_dim, 64, 3, 1, 1), nn.ReLU(), nn.Conv2d(64, 64, 3, 1, 1), nn.ReLU(),
                  nn.Conv2d(64, 64, 3, 1, 1), nn.ReLU(), nn.Conv2d(64, f_dim, 5, 1, 2)]
        if not no_BN:
            layers.append(nn.BatchNorm2d(f_dim))
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        mean, std = x.new_tensor(self.mean), x.new_tensor(self.std)
        x[:, :3] = (x[:, :3] - mean[:, None, None]) / std[:, None, None]
        y = self.net(x)
        return x[:, 3:] + y if self.fusion_residule else y


class ExposureMLP(nn.Module):
    """The exposure network of the coarse model (nerfh_nff.py:511-522): tiny-cuda-nn `FullyFusedMLP`, 10 -> 32 -> 32 -> 32
    -> 12, ReLU, no biases, restated in plain torch -- a caller-side module after the render path.
    tiny-cuda-nn is an un-vendored dependency with no pinned version (reference README.md:24) and no reference test holds
    a vector for it: PARITY UNPINNED.  Restated from its published algorithm: one flat `params` buffer holding the weight
    matrices [out, in] row-major in layer order, input width padded to a multiple of 16 (10 -> 16), output width padded
    to 16 (12 -> 16), Xavier-uniform init; the arithmetic here is fp32 (tiny-cuda-nn: fp16 operands, fp32 accumulate)."""
    SHAPES = ((32, 16), (32, 32), (32, 32), (16, 32))

    def __init__(self, n_in=10, n_out=12):
        super().__init__()
        self.n_in, self.n_out = n_in, n_out
        parts = []
        for o, i in self.SHAPES:
            bound = (6.0 / (o + i)) ** 0.5
            parts.append((torch.rand(o * i) * 2 - 1) * bound)
        self.params = nn.Parameter(torch.cat(parts))

    def forward(self, x):
        # tiny-cuda-nn wraps the MLP in an Identity encoding aligned to 16 inputs whose padded dimensions are ONES (so the
        # first layer's weight columns 10..15 act as a learned bias), not zeros
        h = torch.nn.functional.pad(x.to(self.params.dtype), (0, self.SHAPES[0][1] - self.n_in), value=1.0)
        off = 0
        for li, (o, i) in enumerate(self.SHAPES):
            w = self.params[off:off + o * i].view(o, i)
            off += o * i
            h = h @ w.t()
            if li + 1 < len(self.SHAPES):
                h = torch.relu(h)
        return h[:, :self.n_out]


class NeRFH_NFF(nn.Module):
    """The NeFeS field (nerfh_nff.py:421-626): xyz PE(63) -> 8x128 ReLU trunk with a skip at layer
    4 -> softplus sigma, 128 'final' -> [final | dir PE(27)] -> 64 -> 131 (rgb 3 + feature 128),
    plus NeRF-W transient heads on the fine net.

    Storage: ONE flat fp32 nn.Parameter (`flat`) laid out by nefes_param_layout; state_dict() /
    load_state_dict() speak the reference's per-layer keys (xyz_encoding_1.0.weight, ...), so
    reference checkpoints load unchanged.  Only the architecture every reference config uses is
    built (D=8, W=128, skips=[4], 63/27 input channels, f_dim=128).
    """

    def __init__(self, typ, D=8, W=256, skips=[4], in_channels_xyz=63, in_channels_dir=27,
                 encode_appearance=False, in_channels_a=48, encode_transient=False, in_channels_t=16,
                 beta_min=0.1, out_ch_size=3, f_dim=FEATURE_DIM, fusion_residule=False, no_BN=False,
                 precision=None):
        super().__init__()
        if (D, W, list(skips), in_channels_xyz, in_channels_dir, out_ch_size, f_dim) != (8, 128, [4], 63, 27, 3, 128):
            raise RuntimeError("nefes_b200: only D=8, W=128, skips=[4], xyz 63, dir 27, rgb 3 + 128 features is built "
                               f"(got D={D} W={W} skips={skips} xyz={in_channels_xyz} dir={in_channels_dir})")
        torch.manual_seed(0)                                     # nerfh_nff.py:446
        self.typ = typ
        self.D, self.W, self.skips = D, W, skips
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        self.encode_appearance = False if typ == "coarse" else encode_appearance
        self.encode_transient = False if typ == "coarse" else encode_transient
        self.beta_min = beta_min
        self.W_features = f_dim
        self.out_ch_size = out_ch_size + f_dim
        self.fusion_residule, self.no_BN = fusion_residule, no_BN
        self.net_id = L.NET_FINE if self.encode_transient else L.NET_COARSE
        self.precision = precision or DEFAULT_PRECISION

        # Same construction order as the reference so the default init is bit-identical.
        init = OrderedDict()

        def lin(name, fan_in, fan_out):
            layer = nn.Linear(fan_in, fan_out)
            init[name] = (layer.weight.detach(), layer.bias.detach())
        for i in range(D):
            lin(f"xyz_encoding_{i + 1}.0", in_channels_xyz if i == 0 else (W + in_channels_xyz if i in skips else W), W)
        lin("xyz_encoding_final", W, W)
        lin("dir_encoding.0", W + in_channels_dir, W // 2)
        lin("static_sigma.0", W, 1)
        lin("static_rgb.0", W // 2, self.out_ch_size)
        if self.encode_transient:
            lin("transient_encoding.0", W + in_channels_dir, W // 2)
            lin("transient_encoding.2", W // 2, W // 2)
            lin("transient_encoding.4", W // 2, W // 2)
            lin("transient_sigma.0", W // 2, 1)
            lin("transient_rgb.0", W // 2, 3)
            lin("transient_beta.0", W // 2, 1)
        rows, n_params = L.layout(self.net_id)
        assert set(r[0] for r in rows) == set(init), "layout / constructor mismatch"
        flat = torch.empty(n_params)
        for name, out_d, in_d, w_off, b_off in rows:
            w, b = init[name]
            assert tuple(w.shape) == (out_d, in_d)
            flat[w_off:w_off + out_d * in_d] = w.reshape(-1)
            flat[b_off:b_off + out_d] = b
        self.flat = nn.Parameter(flat)
        self._rows = rows
        if typ == "coarse":
            self.fusion_net = FusionNet(self.W_features, fusion_residule, no_BN)
            self.exposure_embedding = ExposureMLP()              # APPLY_HISTOGRAM = True (nerfh_nff.py:20, :511)
        self.sigmoid = nn.Sigmoid()

    # ---- reference-keyed views -------------------------------------------------------------
    def layer_views(self, tensor=None):
        """OrderedDict name.weight / name.bias -> view into `tensor` (default: self.flat.data),
        in the reference's state_dict order."""
        t = self.flat.data if tensor is None else tensor
        by = {r[0]: r for r in self._rows}
        order = [f"xyz_encoding_{i + 1}.0" for i in range(8)] + ["xyz_encoding_final", "dir_encoding.0",
                                                                 "static_sigma.0", "static_rgb.0"]
        if self.net_id == L.NET_FINE:
            order += ["transient_encoding.0", "transient_encoding.2", "transient_encoding.4", "transient_sigma.0",
                      "transient_rgb.0", "transient_beta.0"]
        out = OrderedDict()
        for name in order:
            _, o, i, w_off, b_off = by[name]
            out[name + ".weight"] = t[w_off:w_off + o * i].view(o, i)
            out[name + ".bias"] = t[b_off:b_off + o]
        return out

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k, v in self.layer_views().items():
            destination[prefix + k] = v if keep_vars else v.detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        mine = self.layer_views()
        with torch.no_grad():
            for k, view in mine.items():
                src = state_dict.get(prefix + k)
                if src is None:
                    missing_keys.append(prefix + k)
                elif tuple(src.shape) != tuple(view.shape):
                    error_msgs.append(f"size mismatch for {prefix + k}: {tuple(src.shape)} vs {tuple(view.shape)}")
                else:
                    view.copy_(src)
        own = {prefix + k for k in mine} | {prefix + "flat"}
        child = tuple(prefix + c + "." for c, _ in self.named_children())
        for k in state_dict:
            if k.startswith(prefix) and k not in own and not k.startswith(child):
                unexpected_keys.append(k)

    # ---- kernels ---------------------------------------------------------------------------
    def query(self, pts, viewdirs, mode):
        """pts [N,S,3], viewdirs [N,3] -> raw [N,S,C] (PE fused)."""
        prec = _PREC[self.precision]
        if mode == L.MODE_FULL and self.net_id != L.NET_FINE:
            raise RuntimeError("nefes_b200: transient output requested from the coarse net")
        return ops.field_query(pts, None if mode == L.MODE_SIGMA else viewdirs, self.flat, self.net_id, mode, prec)

    def forward(self, x, sigma_only=False, output_transient=True):
        """Drop-in for nerfh_nff.py:525-576.  `x` is the embedded input [B, 63] / [B, 90]; the raw
        xyz / direction are its leading 3 channels of each block (include_input=True), and the
        encoding is recomputed inside the kernel."""
        xyz = x[:, None, 0:3]
        if sigma_only:
            return self.query(xyz, None, L.MODE_SIGMA)[:, 0]
        dirs = x[:, self.in_channels_xyz:self.in_channels_xyz + 3]
        mode = L.MODE_FULL if (output_transient and self.net_id == L.NET_FINE) else L.MODE_STATIC
        return self.query(xyz, dirs, mode)[:, 0]

    def run_fusion_net(self, rgb, feature, H, W, B):
        """nerfh_nff.py:578-603: rgb [B*H*W,3], feature [B*H*W,128] -> (render_rgb [B,3,H,W], render_feature [B,128,H,W],
        feature_output [B,128,H,W]).  On CUDA the four convolutions + BatchNorm run as engine launches on the pixel-major
        tensors the render produced (no NCHW shuffle on the way in); the NCHW results are views."""
        render_rgb = rgb.reshape(B, H, W, 3).permute(0, 3, 1, 2)
        render_feature = feature.reshape(B, H, W, self.W_features).permute(0, 3, 1, 2)
        if rgb.is_cuda:
            from . import fusion
            out = fusion.fusion_net(self.fusion_net, rgb.reshape(-1, 3), feature.reshape(-1, self.W_features), B, H, W)
            return render_rgb, render_feature, out.reshape(B, H, W, self.W_features).permute(0, 3, 1, 2)
        fusion_input = torch.cat([render_rgb, render_feature], dim=1)
        return render_rgb, render_feature, self.fusion_net(fusion_input)

    def affine_color_transform(self, args, rgb, hist, batch_size):
        """nerfh_nff.py:605-626: rgb [B*N,3], hist [B,10] -> sigmoid(K_b rgb + bias_b), (K_b, bias_b) from the exposure MLP
        of image b's histogram (cast to integers first, as the reference does)."""
        if not (getattr(args, "encode_hist", False) and self.typ == "coarse"):
            raise RuntimeError("nefes_b200: affine_color_transform needs args.encode_hist and the coarse model")
        if rgb.is_cuda:
            from . import fusion
            out, self.a_embedded = fusion.affine_color(self.exposure_embedding.params, rgb, hist, batch_size)
            return out
        self.a_embedded = self.exposure_embedding(hist.long()).float()
        kernel = self.a_embedded[:, :9].reshape(-1, 3, 3)
        bias = self.a_embedded[:, 9:].reshape(-1, 3, 1)
        rgb = rgb.reshape(batch_size, -1, 3)
        rgb = torch.bmm(kernel, rgb.transpose(1, 2)) + bias
        return self.sigmoid(rgb.transpose(1, 2).reshape(-1, 3))


# ------------------------------------------------------------------------------------------------
class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.9,0.999)) semantics (nerfh_nff.py:682) as one fused kernel per
    flat parameter buffer.  param_groups[...]['lr'] can be rewritten by the caller's schedule
    (run_nefes.py:266-270).  grad_scale multiplies gradients first (1/world_size after all-reduce)."""

    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def set_lr(self, lr, group=None):
        """Write a new learning rate into param_groups AND the device scalar the kernel reads.  A step replayed from a
        CUDA graph never passes through step() on the host, so the reference's per-step decay (run_nefes.py:266-270)
        must come through here (outside capture) when the step is graph-replayed."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("nefes_b200.FlatAdam: change the learning rate outside graph capture")
        for gi, grp in enumerate(self.param_groups):
            if group is not None and gi != group:
                continue
            grp["lr"] = float(lr)
            for p in grp["params"]:
                st = self.state.get(p)
                if st:
                    st["dev"][1] = float(lr)
                    st["lr_host"] = float(lr)

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("nefes_b200.FlatAdam: parameters must be contiguous fp32 CUDA tensors")
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    # (step count, lr) live on the device so that a captured CUDA graph of the step replays correctly
                    st["dev"] = torch.tensor([0.0, float(group["lr"])], device=p.device)
                    st["lr_host"] = float(group["lr"])
                capturing = torch.cuda.is_current_stream_capturing()
                if float(group["lr"]) != st["lr_host"]:
                    if capturing:
                        raise RuntimeError("nefes_b200.FlatAdam: change the learning rate outside graph capture")
                    st["dev"][1] = float(group["lr"])
                    st["lr_host"] = float(group["lr"])
                st["step"] += 1                                  # host mirror (replays of a graph do not pass here)
                ops.adam_step_dev(p, p.grad, st["exp_avg"], st["exp_avg_sq"], st["dev"], group["betas"][0],
                                  group["betas"][1], group["eps"], grad_scale)


def create_nerf(args, device=None):
    """Drop-in for nerfh_nff.py:628-737: (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    device = torch.device(device or "cuda")
    embed_fn, input_ch, _ = get_embedder(args.multires, args.i_embed, getattr(args, "reduce_embedding", -1))
    embeddirs_fn, input_ch_views = None, 0
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views, _ = get_embedder(args.multires_views, args.i_embed,
                                                       getattr(args, "reduce_embedding", -1))
    skips = [4]
    # The factory the reference's scripts enter through builds the FAST path: bf16 tcgen05 layer chains (1.9 M rays/s; stated
    # tolerances in DESIGN.md section 3, refined poses within 1 mm / 0.01 deg of the fp32 reference).  args.nefes_precision or
    # NEFES_PRECISION select "tf32" (tensor cores at <= 1e-3 of the reference, 0.2 M rays/s) or "fp32" (exact SIMT path).
    precision = getattr(args, "nefes_precision", None) or os.environ.get("NEFES_PRECISION", "bf16")
    model = NeRFH_NFF("coarse", D=args.netdepth, W=args.netwidth, skips=skips, in_channels_xyz=input_ch,
                      in_channels_dir=input_ch_views, fusion_residule=getattr(args, "use_fusion_res", False),
                      no_BN=getattr(args, "no_fusion_BN", False), precision=precision).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRFH_NFF("fine", D=args.netdepth, W=args.netwidth, skips=skips, in_channels_xyz=input_ch,
                               in_channels_dir=input_ch_views, encode_appearance=True, encode_transient=True,
                               in_channels_a=getattr(args, "in_channels_a", 48),
                               in_channels_t=getattr(args, "in_channels_t", 16), precision=precision).to(device)
        grad_vars += list(model_fine.parameters())

    network_query_fn = StandardQuery(args.netchunk, embed_fn, embeddirs_fn)
    if getattr(args, "no_grad_update", False):
        grad_vars, optimizer = None, None
    else:
        optimizer = FlatAdam(grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    ckpts = []
    if getattr(args, "ft_path", None) not in (None, "None"):
        ckpts = [args.ft_path]
    elif getattr(args, "basedir", None) and os.path.isdir(os.path.join(args.basedir, args.expname)):
        d = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if "tar" in f]
    if ckpts and not getattr(args, "no_reload", False):
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt["global_step"]
        model.load_state_dict(ckpt["network_fn_state_dict"], strict=False)
        if model_fine is not None:
            model_fine.load_state_dict(ckpt["network_fine_state_dict"])

    render_kwargs_train = dict(network_query_fn=network_query_fn, perturb=args.perturb, N_importance=args.N_importance,
                               N_samples=args.N_samples, network_fn=model, use_viewdirs=args.use_viewdirs,
                               white_bkgd=args.white_bkgd, raw_noise_std=args.raw_noise_std, test_time=False, args=args)
    if model_fine is not None:
        render_kwargs_train["network_fine"] = model_fine
    if args.dataset_type != "llff" or args.no_ndc:
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test.update(perturb=False, raw_noise_std=0., test_time=True)
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer
