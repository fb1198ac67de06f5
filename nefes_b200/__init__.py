"""nefes_b200 -- B200-native NeFeS render engine (sm_100a CUDA behind a C ABI).

Drop-in for the render hot path of ActiveVisionLab/NeFeS: `render`, `render_rays`, `sample_pdf`
(script/models/rendering.py), `get_rays`, `get_rays_batch` (script/models/ray_utils.py),
`raw2outputs_NeRFH_NFF`, `run_network_NeRFH_NFF`, `NeRFH_NFF`, `create_nerf`
(script/models/nerfh_nff.py).  Importing the package does not load the CUDA library; the first
kernel call does, and raises if it is missing (there is no CPU fallback)."""
from .nerfh_nff import (FlatAdam, FusionNet, NeRFH_NFF, create_nerf, get_embedder, img2mse, mse2psnr,  # noqa: F401
                        raw2outputs_NeRFH_NFF, run_network_NeRFH_NFF, StandardQuery, to8b)
from .batching import gather_ray_batch, select_random_patches, select_random_pixels  # noqa: F401
from .losses import NerfWLoss, ColorFeatureFusionNerfWLoss  # noqa: F401
from .ray_utils import get_rays, get_rays_batch  # noqa: F401
from .rendering import batchify_rays, render, render_rays, sample_pdf  # noqa: F401

__version__ = "0.1.0"
