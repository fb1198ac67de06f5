"""Import the unmodified reference modules (from /root/reference in the build container, or from the staged copy
oracle/_ref/ on the GPU box).  TEST INFRASTRUCTURE: used by oracle/make_golden.py, tests and bench.py --impl reference.

Four third-party modules the reference imports at module level but never executes on this path are stubbed
(imageio, matplotlib, tinycudann, lietorch) -- SURVEY.md section 8c."""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))


def reference_root():
    for r in CANDIDATES:
        if os.path.isfile(os.path.join(r, "script", "models", "rendering.py")):
            return r
    return None


def import_reference(root=None):
    """-> (models.rendering, models.nerfh_nff, models.ray_utils) of the reference, or raises ImportError."""
    root = root or reference_root()
    if root is None:
        raise ImportError("no reference tree (neither /root/reference nor oracle/_ref; run oracle/build_ref.py in the build container)")
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "lietorch"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["lietorch"], "SE3"):
        sys.modules["lietorch"].SE3 = None
    if "tinycudann" not in sys.modules:
        tcnn = types.ModuleType("tinycudann")

        class _Placeholder(torch.nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
        tcnn.Network = _Placeholder
        tcnn.Encoding = _Placeholder
        sys.modules["tinycudann"] = tcnn
    for p in (os.path.join(root, "script"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import models.rendering as R
    import models.nerfh_nff as M
    import models.ray_utils as U
    return R, M, U


def reference_render_kwargs(M, netchunk=1 << 21):
    """The two fields and the render_kwargs dict that create_nerf builds (nerfh_nff.py:628-737), without its hard-coded
    cuda device and log-directory scan: constructors, embedders and the network_query_fn lambda (:667-675) as written there."""
    kw = dict(D=8, W=128, skips=[4], in_channels_xyz=63, in_channels_dir=27)
    coarse = M.NeRFH_NFF("coarse", **kw)
    fine = M.NeRFH_NFF("fine", encode_appearance=True, encode_transient=True, **kw)
    embed_fn, _, _ = M.get_embedder(10, 0, -1)
    embeddirs_fn, _, _ = M.get_embedder(4, 0, -1)

    class Args:
        nerfh_nff, use_fine_only, NeRFW, transient_at_test = True, False, True, True
    query = lambda inputs, viewdirs, ts, fn, typ, output_transient, test_time, store_rgb: \
        M.run_network_NeRFH_NFF(inputs, viewdirs, ts, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, typ=typ,
                                output_transient=output_transient, netchunk=netchunk, test_time=test_time, store_rgb=store_rgb)
    base = dict(network_query_fn=query, N_importance=64, N_samples=64, network_fn=coarse, network_fine=fine,
                use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False)
    return coarse, fine, base
