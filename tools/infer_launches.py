"""One inference render (test_time, no gradients) of 32768 rays in one chunk, three times: run under
`ncu --metrics gpu__time_duration.sum` to list the launches of the C5 sweep's inner step.  Usage: python tools/infer_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nefes_b200 as nb
dev = torch.device("cuda")
class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 21
c = nb.NeRFH_NFF("coarse", W=128, precision="bf16").to(dev)
f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True, precision="bf16").to(dev)
kw = dict(network_query_fn=nb.StandardQuery(Args.netchunk), N_importance=64, N_samples=64, network_fn=c, network_fine=f, use_viewdirs=True,
          white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=10., perturb=0., raw_noise_std=0., test_time=True)
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
