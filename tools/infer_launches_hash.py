"""Front-end B (HashGrid + SH, nerfh_tcnn) inference render of 32768 rays, three times: run under
`ncu --metrics gpu__time_duration.sum` to list the launches.  Usage: python tools/infer_launches_hash.py [log2T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nefes_b200 as nb
from nefes_b200 import _lib
from nefes_b200.hashgrid import NeRFH_TCNN, TcnnQuery
dev = torch.device("cuda")
log2T = int(sys.argv[1]) if len(sys.argv) > 1 else 19
hc = NeRFH_TCNN("coarse", bound=25, log2_hashmap_size=log2T).to(dev)
hf = NeRFH_TCNN("fine", encode_appearance=True, encode_transient=True, in_channels_a=50, in_channels_t=20, bound=25, log2_hashmap_size=log2T).to(dev)
class HArgs:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test = False, False, True, True
kw = dict(network_query_fn=TcnnQuery(1 << 21), N_importance=64, N_samples=64, network_fn=hc, network_fine=hf, use_viewdirs=True,
          white_bkgd=False, args=HArgs(), ndc=False, lindisp=False, near=0., far=10., perturb=0., raw_noise_std=0., test_time=True)
_lib.lib().nefes_gemm_mode(1)
n = 32768
g = torch.Generator(device=dev).manual_seed(0)
ro = torch.randn(n, 3, device=dev, generator=g) * 0.1
rd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1)
hist = torch.zeros(1, 10, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for rep in range(3):
        torch.cuda.synchronize(); ev0.record()
        nb.render(60, 106, 93.0, chunk=32768, rays=(ro, rd), img_idx=hist, **kw)
        ev1.record(); torch.cuda.synchronize()
        print(f"render {rep}: {ev0.elapsed_time(ev1):.3f} ms, {n / ev0.elapsed_time(ev1) / 1e3:.2f} M rays/s", flush=True)
