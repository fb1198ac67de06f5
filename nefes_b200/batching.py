"""Ray-batch assembly in front of `render`: the selection / gather halves of
script/run_nefes.py:42-76 (`render_nerf_random_ray`) and :78-107 (`render_nerf_random_patch`), without the
per-image host loops -- one device-side draw, one gather per tensor.  The render call itself stays the caller's:

    sel = select_random_pixels(B, H, W, N_rand, valid_inds)                 # [B, N_rand] flat pixel indices
    rays, target_s, target_f, hist = gather_ray_batch(H, W, focal, pose, sel, target, feature_target, hist)
    rgb, disp, acc, extras = render(H, W, focal, chunk=args.chunk, rays=rays, retraw=True, img_idx=hist, **kw)
"""
import torch

from .ray_utils import get_rays_batch


def select_random_pixels(B, H, W, N_rand, valid_inds=None, device=None, generator=None):
    """N_rand distinct flat pixel indices (row-major, j*W + i) per image, uniformly at random -- the reference's
    `np.random.choice(n, size=[N_rand], replace=False)` per image (run_nefes.py:51-65), drawn on the device as the
    first N_rand entries of an argsort of uniform noise.  valid_inds: per-image index tensors of the static pixels
    (`args.semantic`, run_nefes.py:130-133); invalid pixels are never selected."""
    n = H * W
    device = torch.device(device) if device is not None else (valid_inds[0].device if valid_inds else torch.device("cpu"))
    keys = torch.rand(B, n, device=device, generator=generator)
    if valid_inds is not None:
        if len(valid_inds) != B:
            raise RuntimeError("nefes_b200: one valid-index tensor per image is required")
        mask = torch.zeros(B, n, dtype=torch.bool, device=device)
        for b, v in enumerate(valid_inds):
            if v.numel() < N_rand:
                raise RuntimeError(f"nefes_b200: image {b} has {v.numel()} valid pixels, fewer than N_rand={N_rand}")
            mask[b, v.to(device)] = True
        keys = torch.where(mask, keys, torch.full_like(keys, 2.0))
    elif N_rand > n:
        raise RuntimeError(f"nefes_b200: N_rand={N_rand} exceeds the {n} pixels of an image")
    return keys.argsort(dim=1)[:, :N_rand]


def select_random_patches(H, W, num_crops=7, crop_size=16, device=None, generator=None):
    """Flat pixel indices [num_crops * crop_size^2] of `num_crops` random crop_size x crop_size patches, the same for
    every image of the batch (run_nefes.py:86-95: top-left corners uniform in [0, H - crop) x [0, W - crop))."""
    device = torch.device(device or "cpu")
    h0 = torch.randint(0, H - crop_size, (num_crops,), device=device, generator=generator)
    w0 = torch.randint(0, W - crop_size, (num_crops,), device=device, generator=generator)
    d = torch.arange(crop_size, device=device)
    rows = (h0[:, None] + d[None, :])[:, :, None]                     # [crops, cs, 1]
    cols = (w0[:, None] + d[None, :])[:, None, :]                     # [crops, 1, cs]
    return (rows * W + cols).reshape(-1)


def gather_ray_batch(H, W, focal, pose, sel, target=None, feature_target=None, hist=None):
    """pose [B,3,4]; sel [B,n] (per image) or [n] (shared, the patch case) flat pixel indices; target [B,H,W,3] and
    feature_target [B,H,W,C] channel-last, hist [B,10].  Returns (batch_rays [2, B*n, 3], target_s [B*n,3],
    target_f [B*n,C], hist [B*n,10]) in the reference's image-major order (run_nefes.py:66-74, :98-106)."""
    B = pose.shape[0]
    if sel.dim() == 1:
        sel = sel[None].expand(B, -1)
    n = sel.shape[1]
    rays_o, rays_d = get_rays_batch(H, W, focal, pose)                # [B,H,W,3]

    def take(x):
        c = x.shape[-1]
        return torch.gather(x.reshape(B, H * W, c), 1, sel[..., None].expand(-1, -1, c)).reshape(B * n, c)
    batch_rays = torch.stack([take(rays_o), take(rays_d)], 0)
    target_s = take(target) if target is not None else None
    target_f = take(feature_target) if feature_target is not None else None
    hist_e = hist[:, None, :].expand(-1, n, -1).reshape(B * n, -1) if hist is not None else None
    return batch_rays, target_s, target_f, hist_e
