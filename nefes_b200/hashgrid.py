"""Encoder front-end B: multiresolution HashGrid + degree-4 spherical harmonics and the
NeRFH_TCNN-shaped field on top of them -- the host-side mirror of script/models/nerfh_tcnn.py:15-284, whose
arithmetic the reference delegates to tiny-cuda-nn (tcnn.Encoding / tcnn.Network).

    HashGridEncoding   tcnn.Encoding(otype="HashGrid", ...)          nerfh_tcnn.py:65-75
    SHEncoding         tcnn.Encoding(otype="SphericalHarmonics", 4)  nerfh_tcnn.py:97-103
    NeRFH_TCNN         nerfh_tcnn.py:15-284 (density / color / forward)

Differences, stated: tables and the small MLPs are fp32 here (tcnn stores fp16); FullyFusedMLP is bias-free and so
are these; parity for this row is unpinned (no tiny-cuda-nn in the container, dead code in the reference)."""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L


class _HashEncode(Function):
    @staticmethod
    def forward(ctx, x, table, layout):
        L.need_cuda(x, table)
        xc, tc = L.f32c(x), table.detach()
        M = xc.shape[0]
        out = torch.empty(M, 2 * layout.n_levels, device=xc.device)
        with torch.cuda.device(xc.device):
            L.check(L.lib().nefes_encode_hash_fwd(L.ptr(xc), L.ptr(tc), M, C.byref(layout), L.ptr(out), L.stream_of(xc)),
                    "nefes_encode_hash_fwd")
        ctx.save_for_backward(xc, tc)
        ctx.layout = layout
        return out

    @staticmethod
    def backward(ctx, g):
        xc, tc = ctx.saved_tensors
        g = L.f32c(g)
        need_x, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_table = torch.zeros_like(tc) if need_t else None
        d_x = torch.empty_like(xc) if need_x else None
        if need_x or need_t:
            with torch.cuda.device(xc.device):
                L.check(L.lib().nefes_encode_hash_bwd(L.ptr(xc), L.ptr(g), L.ptr(tc), xc.shape[0], C.byref(ctx.layout),
                                                      L.ptr(d_table), L.ptr(d_x), L.stream_of(xc)), "nefes_encode_hash_bwd")
        return d_x, d_table, None


class _SHEncode(Function):
    @staticmethod
    def forward(ctx, d):
        L.need_cuda(d)
        dc = L.f32c(d)
        out = torch.empty(dc.shape[0], 16, device=dc.device)
        with torch.cuda.device(dc.device):
            L.check(L.lib().nefes_encode_sh_fwd(L.ptr(dc), dc.shape[0], L.ptr(out), L.stream_of(dc)), "nefes_encode_sh_fwd")
        ctx.save_for_backward(dc)
        return out

    @staticmethod
    def backward(ctx, g):
        (dc,) = ctx.saved_tensors
        g = L.f32c(g)
        dd = torch.empty_like(dc)
        with torch.cuda.device(dc.device):
            L.check(L.lib().nefes_encode_sh_bwd(L.ptr(dc), L.ptr(g), dc.shape[0], L.ptr(dd), L.stream_of(dc)), "nefes_encode_sh_bwd")
        return dd


class _Linear(Function):
    """y = act(x W^T) on the engine's fp32 GEMM (bias-free, like tcnn's FullyFusedMLP).  act: 0 none, 1 relu."""

    @staticmethod
    def forward(ctx, x, W, act):
        L.need_cuda(x, W)
        xc, Wc = L.f32c(x), W.detach()
        M, K = xc.shape
        N = Wc.shape[0]
        if Wc.shape[1] != K:
            raise RuntimeError(f"nefes_b200: linear layer expects {Wc.shape[1]} input channels, got {K} (front-end B: in_channels_a / "
                               "in_channels_t must equal 10 histogram bins x the embedding width: 50 / 20)")
        y = torch.empty(M, N, device=xc.device)
        with torch.cuda.device(xc.device):
            L.check(L.lib().nefes_linear_fwd(L.ptr(xc), K, L.ptr(Wc), None, L.ptr(y), N, M, N, K, act, L.stream_of(xc)),
                    "nefes_linear_fwd")
        ctx.save_for_backward(xc, Wc, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, g):
        xc, Wc, y = ctx.saved_tensors
        g = L.f32c(g)
        M, K = xc.shape
        N = Wc.shape[0]
        st = L.stream_of(xc)
        with torch.cuda.device(xc.device):
            if ctx.act == 1:                               # ReLU: gradient w.r.t. the pre-activation
                g = g * (y > 0)
            dW = dx = None
            if ctx.needs_input_grad[1]:
                dW = torch.zeros_like(Wc)
                L.check(L.lib().nefes_linear_wgrad(L.ptr(g), N, L.ptr(xc), K, L.ptr(dW), M, N, K, st), "nefes_linear_wgrad")
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(xc)
                L.check(L.lib().nefes_linear_dgrad(L.ptr(g), N, L.ptr(Wc), L.ptr(dx), K, M, N, K, None, 0, st), "nefes_linear_dgrad")
        return dx, dW, None


def linear(x, W, act=0):
    return _Linear.apply(x, W, int(act))


class HashGridEncoding(nn.Module):
    """tcnn.Encoding(n_input_dims=3, {"otype": "HashGrid", n_levels, n_features_per_level=2, log2_hashmap_size,
    base_resolution, per_level_scale}).  `params` is the flat table, levels concatenated, init U(-1e-4, 1e-4)."""

    def __init__(self, n_levels=16, log2_hashmap_size=19, base_resolution=16, per_level_scale=None, max_resolution=2048):
        super().__init__()
        if per_level_scale is None:
            per_level_scale = math.exp(math.log(max_resolution / base_resolution) / (n_levels - 1))
        self.layout = L.HashLayout()
        L.check(L.lib().nefes_hash_layout(n_levels, log2_hashmap_size, base_resolution, float(per_level_scale),
                                          C.byref(self.layout)), "nefes_hash_layout")
        self.n_output_dims = 2 * n_levels
        self.params = nn.Parameter((torch.rand(int(self.layout.n_entries) * 2) * 2 - 1) * 1e-4)

    def forward(self, x):
        return _HashEncode.apply(x, self.params, self.layout)


class SHEncoding(nn.Module):
    n_output_dims = 16

    def forward(self, d01):
        return _SHEncode.apply(d01)


class NeRFH_TCNN(nn.Module):
    """nerfh_tcnn.py:15-284.  Same constructor arguments and forward(x, d, ts, sigma_only, output_transient)
    contract: [B,1] sigma / [B,4] (rgb, sigma) / [B,9] with the NeRF-W transient head."""

    def __init__(self, typ, W=64, N_vocab=1000, hash_level=16, encode_appearance=False, in_channels_a=48,
                 encode_transient=False, in_channels_t=16, beta_min=0.1, bound=25, log2_hashmap_size=19):
        super().__init__()
        torch.manual_seed(0)                                       # nerfh_tcnn.py:39
        self.typ, self.W, self.bound, self.beta_min = typ, W, bound, beta_min
        self.encode_appearance = False if typ == "coarse" else encode_appearance
        self.in_channels_a = in_channels_a if encode_appearance else 0
        self.encode_transient = False if typ == "coarse" else encode_transient
        self.in_channels_t = in_channels_t
        self.encoder = HashGridEncoding(hash_level, log2_hashmap_size, 16, None, 2048)
        self.encoder_dir = SHEncoding()

        def mlp(dims):                                             # bias-free, tcnn-style uniform init
            return nn.ParameterList([nn.Parameter((torch.rand(o, i) * 2 - 1) * math.sqrt(6.0 / (i + o)))
                                     for i, o in zip(dims[:-1], dims[1:])])
        self.sigma_net = mlp([self.encoder.n_output_dims, 64, W + 1])                 # :79-89
        self.embedding_a = nn.Embedding(N_vocab, 5)
        in_color = W + 16 + (self.in_channels_a if self.encode_appearance else 0)
        self.color_net = mlp([in_color, 64, 64, 3])                                    # :111-121
        if self.encode_transient:
            self.embedding_t = nn.Embedding(N_vocab, 2)
            self.transient_color_net = mlp([W + 16 + in_channels_t, 64, 64, 64, 5])    # :129-139

    @staticmethod
    def _run(net, h):
        for i, Wt in enumerate(net):
            h = linear(h, Wt, 1 if i + 1 < len(net) else 0)
        return h

    def input_norm(self, x):
        return (x + self.bound) / (2 * self.bound)

    def density(self, x, norm_input=True):
        if norm_input:
            x = self.input_norm(x)
        h = self._run(self.sigma_net, self.encoder(x))
        return {"sigma": torch.relu(h[..., 0]), "geo_feat": h[..., 1:]}

    def color(self, x, d, ts=None, mask=None, geo_feat=None, transient=False, norm_input=True):
        d = self.encoder_dir((d + 1) / 2)
        parts = [d, geo_feat]
        if self.encode_appearance:
            parts.append(self.embedding_a(ts.long()).reshape(ts.shape[0], -1))
        rgbs = torch.sigmoid(self._run(self.color_net, torch.cat(parts, -1)))
        if transient:
            t = torch.cat([d, geo_feat, self.embedding_t(ts.long()).reshape(ts.shape[0], -1)], -1)
            t = self._run(self.transient_color_net, t)
            return torch.cat([rgbs, torch.sigmoid(t[..., 1:4]), torch.relu(t[..., 0:1]), torch.relu(t[..., 4:5])], 1)
        return rgbs

    def forward(self, x, d, ts=None, sigma_only=False, output_transient=False):
        dens = self.density(x)
        sigma = dens["sigma"]
        if sigma_only:
            return sigma[..., None]
        rgbs = self.color(x, d, ts=ts, geo_feat=dens["geo_feat"], transient=output_transient)
        if not output_transient:
            return torch.cat([rgbs, sigma[..., None]], 1)
        return torch.cat([rgbs[..., :3], sigma[..., None], rgbs[..., 3:]], 1)


def run_NeRFH_TCNN(inputs, viewdirs, ts, fn, typ, output_transient, netchunk=1024 * 64, test_time=False, store_rgb=False):
    """Drop-in for nerfh_tcnn.py:368-440, the `network_query_fn` of front-end B (coarse = NeRF, fine = NeRF-W): inputs
    [N_rays, N_samples, 3], viewdirs [N_rays, 3], ts [N_rays, 10] (the image's histogram, indexes the appearance / transient
    embeddings) -> raw [N_rays, N_samples, 1] (coarse at test time), [.., 4] (rgb, sigma) or [.., 9] (+ transient rgb, sigma,
    beta), in netchunk-sized pieces like the reference."""
    n_rays, n_samples = inputs.shape[0], inputs.shape[1]
    flat = inputs.reshape(-1, 3)
    if typ == 'coarse' and test_time and not store_rgb:
        out = torch.cat([fn(flat[i:i + netchunk], None, sigma_only=True) for i in range(0, flat.shape[0], netchunk)], 0)
        return out.reshape(n_rays, n_samples, -1)
    dirs = viewdirs[:, None].expand(inputs.shape).reshape(-1, 3)
    if typ == 'coarse':
        out = torch.cat([fn(flat[i:i + netchunk], dirs[i:i + netchunk], output_transient=output_transient)
                         for i in range(0, flat.shape[0], netchunk)], 0)
        return out.reshape(n_rays, n_samples, -1)
    if n_samples == 1 and test_time:
        ts_pt = ts[0:1].expand(n_rays, -1)
    else:
        ts_pt = ts[:, None, :].expand(-1, n_samples, -1).reshape(-1, ts.shape[-1])     # repeat 'n1 c -> (n1 n2) c'
    out = torch.cat([fn(flat[i:i + netchunk], dirs[i:i + netchunk], ts=ts_pt[i:i + netchunk], output_transient=output_transient)
                     for i in range(0, flat.shape[0], netchunk)], 0)
    return out.reshape(n_rays, n_samples, -1)


class TcnnQuery:
    """The `network_query_fn` closure the tcnn create_nerf builds (nerfh_tcnn.py:330-340) as an object.  render_rays takes
    the staged route with it (sampling, compositing and sample_pdf on the engine; the hash / SH gathers and the small heads
    are nefes_encode_hash_*, nefes_encode_sh_* and the engine GEMM)."""

    def __init__(self, netchunk=1024 * 64):
        self.netchunk = int(netchunk)

    def __call__(self, inputs, viewdirs, ts, network_fn, typ, output_transient, test_time, store_rgb):
        return run_NeRFH_TCNN(inputs, viewdirs, ts, network_fn, typ, output_transient, netchunk=self.netchunk,
                              test_time=test_time, store_rgb=store_rgb)
