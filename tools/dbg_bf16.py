import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nefes_oracle as O
import nefes_b200 as nb
from nefes_b200 import _lib as L, ops
DEV = "cuda"
w = np.load("tests/golden/weights.npz")
wc = {k[7:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("coarse/")}
wf = {k[5:]: torch.from_numpy(w[k]) for k in w.files if k.startswith("fine/")}
c = nb.NeRFH_NFF("coarse", W=128); f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True)
c.load_state_dict(wc, strict=False); f.load_state_dict(wf); c.to(DEV); f.to(DEV)
def nrm(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
gen = torch.Generator().manual_seed(20)
n, s = int(os.environ.get("NR", 37)), 64
pts = torch.rand(n, s, 3, generator=gen) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
for model, P, mode, typ, tr in ((f, wf, 2, "fine", True), (c, wc, 1, "coarse", False), (c, wc, 0, "coarse", False)):
    Pg = O.clone_params(P, requires_grad=True)
    ref = O.query_field(Pg, pts, dirs, typ, tr, test_time=(mode == 0))
    k = torch.randn(ref.shape, generator=gen)
    (ref * k).sum().backward()
    for prec in (L.PREC_FP32, L.PREC_BF16):
        model.zero_grad()
        raw = ops.field_query(pts.to(DEV), None if mode == 0 else dirs.to(DEV), model.flat, model.net_id, mode, prec)
        (raw * k.to(DEV)).sum().backward()
        views = model.layer_views(model.flat.grad)
        print(f"mode {mode} prec {prec}: raw nrm err {nrm(raw, ref):.3e}; per-column-block:",
              " ".join(f"{nrm(raw[..., a:b], ref[..., a:b]):.1e}" for a, b in ((0, 3), (3, 131), (131, 132), (132, 135), (135, 137)) if b <= ref.shape[-1]))
        for key, rg in Pg.items():
            if rg.grad is not None:
                print(f"    {key:32s} {nrm(views[key], rg.grad):.3e}")
