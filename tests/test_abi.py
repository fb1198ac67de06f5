"""CPU: the C-ABI library loads and exports every symbol include/nefes_b200.h declares; the flat
parameter layout matches the reference's state_dict; host-side argument checks fire without a GPU."""
import ctypes
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nefes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nefes_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nefes_b200 import _lib
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == syms, "ctypes table and header disagree"
    assert lib.nefes_version() == 100


def test_layout_matches_reference_state_dict(weights):
    from nefes_b200 import _lib
    wc, wf = weights
    for net, ref in ((0, wc), (1, wf)):
        rows, n = _lib.layout(net)
        assert n == sum(v.numel() for v in ref.values())
        spans = []
        for name, o, i, w_off, b_off in rows:
            assert tuple(ref[name + ".weight"].shape) == (o, i)
            assert tuple(ref[name + ".bias"].shape) == (o,)
            spans += [(w_off, w_off + o * i), (b_off, b_off + o)]
        spans.sort()
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])), "layout has gaps or overlaps"


def test_host_argument_checks_without_gpu():
    from nefes_b200 import _lib
    lib = _lib.lib()
    assert lib.nefes_get_rays_fwd(None, 1, 60, 80, 65.0, None, None, None) == 1          # NEFES_EINVAL
    assert b"null" in lib.nefes_last_error()
    a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    assert lib.nefes_mlp_workspace(0, 2, 0, 64, 1, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 1
    assert b"fine net" in lib.nefes_last_error()
    assert lib.nefes_mlp_workspace(1, 2, 0, 128 * 64, 64, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 0
    assert a.value > 0 and c.value > 0


def test_model_state_dict_round_trip(weights):
    import nefes_b200 as nb
    wc, wf = weights
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True)
    sd = f.state_dict()
    assert list(sd) == list(wf)                      # same keys, same order as the reference
    assert all(torch.equal(sd[k], wf[k]) for k in wf)   # default init == reference constructor init
    c = nb.NeRFH_NFF("coarse", W=128)
    sdc = c.state_dict()
    assert all(torch.equal(sdc[k], wc[k]) for k in wc)
    assert any(k.startswith("fusion_net.net.") for k in sdc)
    # the coarse model also carries the reference's caller-side modules under the reference's keys
    assert sdc["exposure_embedding.params"].shape == (3072,)         # tcnn FullyFusedMLP 10(->16)x32, 32x32 x2, 32x12(->16)
    missing = c.load_state_dict({**wc, "exposure_embedding.params": torch.zeros(3072)}, strict=False)
    assert all(k.startswith("fusion_net") for k in missing.missing_keys) and not missing.unexpected_keys
    assert float(c.exposure_embedding.params.detach().abs().max()) == 0.0


def test_affine_color_transform_matches_its_definition():
    """nerfh_nff.py:605-626 restated on the torch exposure MLP (tiny-cuda-nn FullyFusedMLP: parity unpinned)."""
    import nefes_b200 as nb
    c = nb.NeRFH_NFF("coarse", W=128)

    class Args:
        encode_hist = True
    g = torch.Generator().manual_seed(4)
    B, n = 3, 7
    rgb = torch.rand(B * n, 3, generator=g, requires_grad=True)
    hist = torch.randint(0, 4, (B, 10), generator=g).float()
    out = c.affine_color_transform(Args(), rgb, hist, B)
    assert out.shape == (B * n, 3) and bool(((out >= 0) & (out <= 1)).all())
    e = c.exposure_embedding
    # tiny-cuda-nn pads the 10 inputs to 16 with ONES (Identity encoding): weight columns 10..15 act as a learned bias
    h = torch.nn.functional.pad(hist, (0, 6), value=1.0)
    assert float(e.params[:32 * 16].view(32, 16)[:, 10:].abs().max()) > 0
    off = 0
    for li, (o, i) in enumerate(e.SHAPES):
        h = h @ e.params[off:off + o * i].view(o, i).t()
        off += o * i
        h = torch.relu(h) if li < 3 else h
    ref = torch.stack([torch.sigmoid(h[b, :9].view(3, 3) @ rgb[b * n + k] + h[b, 9:12]) for b in range(B) for k in range(n)])
    assert float((out - ref).abs().max()) < 1e-6
    out.sum().backward()
    assert rgb.grad is not None and e.params.grad is not None and float(e.params.grad.abs().sum()) > 0
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True)
    import pytest
    with pytest.raises(RuntimeError, match="coarse"):
        f.affine_color_transform(Args(), rgb, hist, B)


def test_cpu_tensors_are_rejected():
    import pytest
    import nefes_b200 as nb
    with pytest.raises(RuntimeError, match="CUDA"):
        nb.get_rays(4, 4, 10.0, torch.eye(4))


def test_render_rays_workspace_is_host_arithmetic():
    """nefes_render_rays_workspace sizes the two workspaces of the one-call path on the host; bad configs are refused."""
    import ctypes as C
    from nefes_b200 import _lib as L
    cfg = L.RenderCfg(64, 64, L.PREC_BF16, 0, 1, 1, L.NET_COARSE, L.NET_FINE, 0.1, 0)
    k, a, b = C.c_int64(), C.c_int64(), C.c_int64()
    assert L.lib().nefes_render_rays_workspace(C.byref(cfg), 6144, C.byref(k), C.byref(a), C.byref(b)) == 0
    # keep >= sample points + both tile-major raw blocks
    assert k.value >= 6144 * (64 + 128) * 12 + 6144 * 64 * 132 * 4 + 6144 * 128 * 137 * 4
    assert a.value > 0 and b.value > 0 and k.value % 256 == 0
    k2 = C.c_int64()
    assert L.lib().nefes_render_rays_workspace(C.byref(cfg), 12288, C.byref(k2), C.byref(a), C.byref(b)) == 0
    assert k2.value > k.value
    bad = L.RenderCfg(64, 300, L.PREC_BF16, 0, 1, 1, L.NET_COARSE, L.NET_FINE, 0.1, 0)
    assert L.lib().nefes_render_rays_workspace(C.byref(bad), 16, C.byref(k), C.byref(a), C.byref(b)) == 1
    assert b"sample counts" in L.lib().nefes_last_error()
    assert L.lib().nefes_render_rays_fwd(C.byref(cfg), None, 16, None, None, None, None) == 1


def test_header_is_plain_c_and_matches_the_ctypes_table():
    """include/nefes_b200.h must compile as C99 (the boundary is a C ABI: no C++ or torch types) and declare exactly the
    entry points the ctypes table binds."""
    import os, re, shutil, subprocess
    from nefes_b200 import _lib as L
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "nefes_b200.h")
    gcc = shutil.which("gcc")
    if gcc:
        r = subprocess.run([gcc, "-x", "c", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(hdr).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(nefes_[a-z0-9_]+)\(", text, flags=re.M))
    assert declared == set(L.exported_symbols()), (declared ^ set(L.exported_symbols()))
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # no torch types in any signature


def test_random_pixel_and_patch_selection():
    """batching.select_random_pixels / select_random_patches (run_nefes.py:51-65, :86-95): distinct, in range, inside
    the valid set; patches are whole crop x crop blocks inside the image."""
    import pytest
    import nefes_b200 as nb
    g = torch.Generator().manual_seed(2)
    H, W, B, n = 12, 20, 3, 50
    sel = nb.select_random_pixels(B, H, W, n, generator=g)
    assert sel.shape == (B, n) and int(sel.min()) >= 0 and int(sel.max()) < H * W
    assert all(len(set(row.tolist())) == n for row in sel)
    valid = [torch.arange(0, H * W, 2), torch.arange(100, 160), torch.arange(H * W)]
    sel = nb.select_random_pixels(B, H, W, n, valid_inds=valid, generator=g)
    for b in range(B):
        assert set(sel[b].tolist()) <= set(valid[b].tolist()) and len(set(sel[b].tolist())) == n
    with pytest.raises(RuntimeError, match="valid pixels"):
        nb.select_random_pixels(B, H, W, 61, valid_inds=valid, generator=g)
    full = nb.select_random_pixels(1, H, W, H * W, generator=g)                      # all pixels = a permutation
    assert sorted(full[0].tolist()) == list(range(H * W))
    pat = nb.select_random_patches(40, 60, num_crops=5, crop_size=8, generator=g)
    assert pat.shape == (5 * 64,)
    for k in range(5):
        blk = pat[k * 64:(k + 1) * 64].reshape(8, 8)
        r, c = blk // 60, blk % 60
        assert bool((r[:, 0] == r[:, -1]).all()) and bool((c[0] == c[-1]).all())
        assert bool((r[1:, 0] - r[:-1, 0] == 1).all()) and bool((c[0, 1:] - c[0, :-1] == 1).all())
        assert int(r.max()) < 40 and int(c.max()) < 60


def _build_c_example(tmp_path):
    import os, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if not gcc or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        return None
    exe = os.path.join(str(tmp_path), "render_c_abi")
    libdir = os.path.join(root, "nefes_b200", "lib")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(root, "examples", "render_c_abi.c"), "-L", libdir, "-lnefes_b200",
                        "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_host_links_against_the_library(tmp_path):
    """examples/render_c_abi.c -- a C99 host with no Python and no torch -- compiles against include/nefes_b200.h and
    links against libnefes_b200.so (running it needs a GPU: tests/test_gpu_tiles.py)."""
    import pytest
    if _build_c_example(tmp_path) is None:
        pytest.skip("gcc or the CUDA runtime headers are not available")
