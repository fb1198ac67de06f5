"""Per-kernel library timing (nefes_prof_*) of the bf16 field forward at the bench shape.  Usage: python tools/prof_fwd.py [rays]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import nefes_b200 as nb
from nefes_b200 import _lib as L, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
g = torch.Generator(device="cuda").manual_seed(0)
for name, S, mode in (("fine", 128, L.MODE_FULL), ("coarse", 64, L.MODE_STATIC), ("sigma", 64, L.MODE_SIGMA)):
    f = (nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True) if name == "fine" else nb.NeRFH_NFF("coarse", W=128)).cuda()
    pts = torch.rand(n, S, 3, device="cuda", generator=g) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    for grad in (True, False):
        f.flat.requires_grad_(grad)
        for i in range(3):
            ops.field_query(pts, dirs if mode != L.MODE_SIGMA else None, f.flat, f.net_id, mode, L.PREC_BF16)
        torch.cuda.synchronize()
        L.lib().nefes_prof_enable(1)
        for i in range(10):
            ops.field_query(pts, dirs if mode != L.MODE_SIGMA else None, f.flat, f.net_id, mode, L.PREC_BF16)
        torch.cuda.synchronize()
        buf = C.create_string_buffer(1 << 16)
        L.lib().nefes_prof_report(buf, 1 << 16)
        L.lib().nefes_prof_enable(0)
        for k, v in json.loads(buf.value.decode()).items():
            ms = v["ms"] / v["launches"]
            print(f"{name:6s} saves={'on ' if grad else 'off'} {k:18s} {ms:.4f} ms  {v['alg_bytes'] / v['launches'] / ms / 1e6:7.0f} GB/s  {v['alg_flops'] / v['launches'] / ms / 1e9:6.0f} TFLOP/s")
