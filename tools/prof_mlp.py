"""Small driver for ncu: fine-net field query forward + backward at the bench shape (6144 rays x 128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nefes_b200 as nb
from nefes_b200 import _lib as L, ops
prec = L.PREC_BF16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else L.PREC_FP32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6144
f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
pts = torch.rand(n, 128, 3, device="cuda", generator=g) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
for _ in range(2):
    f.zero_grad()
    raw = ops.field_query(pts, dirs, f.flat, f.net_id, L.MODE_FULL, prec)
    raw.backward(torch.randn_like(raw))
torch.cuda.synchronize()
print("ok", float(f.flat.grad.abs().sum()))
