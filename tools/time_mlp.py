"""CUDA-event timing of the field MLP (K5) forward and backward at the bench shape: fine FULL (6144 x 128) and
coarse STATIC (6144 x 64).  Usage: python tools/time_mlp.py [bf16|fp32] [rays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nefes_b200 as nb
from nefes_b200 import _lib as L, ops
prec = L.PREC_BF16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else L.PREC_FP32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6144
g = torch.Generator(device="cuda").manual_seed(0)
FLOP = {"fine": 184064 * 2, "coarse": 165632 * 2}
for name, S, mode in (("fine", 128, L.MODE_FULL), ("coarse", 64, L.MODE_STATIC)):
    f = (nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True) if name == "fine" else nb.NeRFH_NFF("coarse", W=128)).cuda()
    pts = torch.rand(n, S, 3, device="cuda", generator=g) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    tf, tb = [], []
    for it in range(8):
        f.zero_grad()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        raw = ops.field_query(pts, dirs, f.flat, f.net_id, mode, prec)
        e[1].record()
        gr = torch.randn_like(raw)
        e[2].record()
        raw.backward(gr)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
    mf, mb = min(tf), min(tb)
    fl = n * S * FLOP[name]
    print(f"{name}: fwd {mf:.3f} ms ({fl / mf / 1e9:.0f} TFLOP/s)  bwd {mb:.3f} ms ({2 * fl / mb / 1e9:.0f} TFLOP/s)  grad|sum| {float(f.flat.grad.abs().sum()):.4f}")
