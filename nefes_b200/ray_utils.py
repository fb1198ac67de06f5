"""Mirror of script/models/ray_utils.py: get_rays (:5-16), get_rays_batch (:46-59) on the K1 kernel.
Differentiable w.r.t. c2w (pose refinement)."""
import torch

from . import ops


def get_rays(H, W, focal, c2w):
    """c2w [3|4,4] -> rays_o, rays_d [H,W,3]."""
    o, d = ops.get_rays(H, W, focal, c2w[None])
    return o[0], d[0]


def get_rays_batch(H, W, focal, c2w):
    """c2w [B,3|4,4] -> rays_o, rays_d [B,H,W,3]."""
    assert c2w.dim() == 3
    return ops.get_rays(H, W, focal, c2w)


def ndc_rays(*a, **k):
    raise RuntimeError("nefes_b200: ndc rays are not on the NeFeS path (ndc=False at nerfh_nff.py:727-730)")
