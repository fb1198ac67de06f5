"""CPU oracle for encoder front-end B (multiresolution HashGrid + spherical harmonics).  TEST
INFRASTRUCTURE ONLY -- imported by tests/ (and bench.py's checker legs), never by nefes_b200/.

PARITY UNPINNED.  The reference reaches this arithmetic only through tiny-cuda-nn
(script/models/nerfh_tcnn.py:65-75 HashGrid, :97-103 SphericalHarmonics), a third-party CUDA dependency
that is neither vendored under /root/reference nor version-pinned (README.md:24 installs git HEAD), in a
model class nothing imports (dead code).  There are no reference tests, golden
vectors or runnable reference outputs for it here.  What follows restates tiny-cuda-nn's published
algorithm (grid.h / spherical_harmonics.h as of the 1.6/1.7 line; Mueller et al. 2022, eq. 2-4):

  level l:  scale_l = 2^(l*log2(b)) * N_min - 1      (fp32),   res_l = ceil(scale_l) + 1
            entries_l = min(next_multiple_of_8(res_l^3), 2^log2_T)
            pos = fma(scale_l, x, 0.5);  cell = floor(pos);  w = pos - cell
            corner index = dense x + y*res + z*res^2 when res^3 <= entries_l, else
                           (x*1 ^ y*2654435761 ^ z*805459861) (uint32), both modulo entries_l
            out[l*F + f] = sum over 8 corners of trilinear weight * table[offset_l + index][f]
  SH degree 4: 16 real basis values of the unit vector 2*d - 1.
"""
from __future__ import annotations

import math

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)


def hash_layout(n_levels=16, log2_T=19, base_res=16, per_level_scale=None, max_res=2048):
    if per_level_scale is None:                      # nerfh_tcnn.py:63: exp(ln(max/min)/(L-1))
        per_level_scale = math.exp(math.log(max_res / base_res) / (n_levels - 1))
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    levels, off = [], 0
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls) * np.float32(base_res) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        n = min(res ** 3, 2 ** 31 - 1)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_T)
        levels.append(dict(scale=float(scale), res=res, size=n, offset=off, dense=res ** 3 <= n))
        off += n
    return levels, off


def hash_encode(x, table, levels, n_feat=2):
    """x [M,3] in [0,1], table [n_entries, n_feat] -> [M, L*n_feat]; differentiable in x and table."""
    outs = []
    for lv in levels:
        scale = lv["scale"]
        pos_exact = (x.detach().double() * scale + 0.5).to(x.dtype)           # fmaf: single rounding
        pos = x * scale + 0.5
        pos = pos + (pos_exact - pos.detach())
        cell = torch.floor(pos.detach())
        w = pos - cell
        cell = cell.long()
        acc = 0
        for corner in range(8):
            bits = [(corner >> d) & 1 for d in range(3)]
            g = cell + torch.tensor(bits)
            wt = 1
            for d in range(3):
                wt = wt * (w[:, d] if bits[d] else 1 - w[:, d])
            if lv["dense"]:
                idx = g[:, 0] + g[:, 1] * lv["res"] + g[:, 2] * lv["res"] ** 2
            else:
                u = [(g[:, d] & 0xFFFFFFFF) for d in range(3)]
                idx = ((u[0] * PRIMES[0]) & 0xFFFFFFFF) ^ ((u[1] * PRIMES[1]) & 0xFFFFFFFF) ^ ((u[2] * PRIMES[2]) & 0xFFFFFFFF)
            idx = idx % lv["size"] + lv["offset"]
            acc = acc + wt[:, None] * table[idx]
        outs.append(acc)
    return torch.cat(outs, -1)


def sh_encode(d01):
    """d01 [M,3] in [0,1] (tcnn convention; the unit vector is 2*d01-1) -> [M,16], degree 4."""
    v = d01 * 2 - 1
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [torch.full_like(x, 0.28209479177387814),
         -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
         1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
         -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
         0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
         0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
         0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
         0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, -1)


def tcnn_field_forward(P, x, d, bound=25.0, levels=None, sigma_only=False):
    """script/models/nerfh_tcnn.py:151-284 (coarse form: no appearance / transient embeddings), fp32, bias-free
    MLPs (FullyFusedMLP has no biases).  P: dict with 'table' [E,2], 'sigma.0' [64,32], 'sigma.1' [65,64],
    'color.0' [64,80], 'color.1' [64,64], 'color.2' [3,64].  Returns [M,4] = (rgb, sigma) or [M,1]."""
    xn = (x + bound) / (2 * bound)                                  # :156
    h = hash_encode(xn, P["table"], levels)
    h = torch.relu(h @ P["sigma.0"].t()) @ P["sigma.1"].t()
    sigma = torch.relu(h[:, 0])                                     # :175
    if sigma_only:
        return sigma[:, None]
    geo = h[:, 1:]
    e = sh_encode((d + 1) / 2)                                      # :209-210
    c = torch.cat([e, geo], -1)                                     # :216
    c = torch.relu(c @ P["color.0"].t())
    c = torch.relu(c @ P["color.1"].t())
    rgb = torch.sigmoid(c @ P["color.2"].t())                      # :220
    return torch.cat([rgb, sigma[:, None]], 1)


def tcnn_field_forward_fine(P, x, d, ts, bound=25.0, levels=None):
    """script/models/nerfh_tcnn.py:151-284, fine form with the NeRF-W heads (`output_transient=True`): appearance embedding
    nn.Embedding(1000, 5) and transient embedding nn.Embedding(1000, 2) indexed by the 10-bin histogram `ts` [M,10] (:107, :125,
    :213, :229) -> 50 / 20 extra inputs; colour net (16 + 64 + 50) -> 64 -> 64 -> 3 sigmoid; transient net (16 + 64 + 20) -> 64 x 3
    -> 5: relu sigma_t (column 0), sigmoid rgb_t (1:4), relu beta (4) (:229-241).  Returns [M,9] = rgb, sigma, rgb_t, sigma_t, beta."""
    xn = (x + bound) / (2 * bound)
    h = hash_encode(xn, P["table"], levels)
    h = torch.relu(h @ P["sigma.0"].t()) @ P["sigma.1"].t()
    sigma = torch.relu(h[:, 0])
    geo = h[:, 1:]
    e = sh_encode((d + 1) / 2)
    a = P["emb_a"][ts.long()].reshape(ts.shape[0], -1)
    c = torch.cat([e, geo, a], -1)
    c = torch.relu(c @ P["color.0"].t())
    c = torch.relu(c @ P["color.1"].t())
    rgb = torch.sigmoid(c @ P["color.2"].t())
    t = torch.cat([e, geo, P["emb_t"][ts.long()].reshape(ts.shape[0], -1)], -1)
    for k in ("trans.0", "trans.1", "trans.2"):
        t = torch.relu(t @ P[k].t())
    t = t @ P["trans.3"].t()
    return torch.cat([rgb, sigma[:, None], torch.sigmoid(t[:, 1:4]), torch.relu(t[:, 0:1]), torch.relu(t[:, 4:5])], 1)
