"""BASELINE config C5-shaped sweep: inference render (test_time, no gradients) of N = 2^12 .. 2^20 synthetic rays through
render() with the reference's chunk of 32768 rays, Cambridge-shaped camera (near 0, far 10), bf16 field path.
Prints rays/s per N (CUDA events, 1 warm-up + 2 timed renders).  Usage: python tools/sweep_rays.py [max_log2=20]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nefes_b200 as nb

dev = torch.device("cuda")
c = nb.NeRFH_NFF("coarse", W=128).to(dev)
f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True).to(dev)
c.precision = f.precision = "bf16"


class Args:
    nerfh_nff, use_fine_only, NeRFW, transient_at_test, netchunk = True, False, True, True, 1 << 22


kw = dict(network_query_fn=nb.StandardQuery(Args.netchunk), N_importance=64, N_samples=64, network_fn=c, network_fine=f,
          use_viewdirs=True, white_bkgd=False, args=Args(), ndc=False, lindisp=False, near=0., far=10., perturb=0.,
          raw_noise_std=0., test_time=True)
g = torch.Generator(device="cuda").manual_seed(0)
hi = int(sys.argv[1]) if len(sys.argv) > 1 else 20
print(f"{'rays':>9s} {'chunks':>6s} {'ms':>9s} {'M rays/s':>9s}")
for lg in range(12, hi + 1):
    n = 1 << lg
    ro = torch.randn(n, 3, device=dev, generator=g) * 0.5
    rd = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1)
    ts = []
    with torch.no_grad():
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rgb, disp, acc, ex = nb.render(60, 106, 93.0, chunk=32768, rays=(ro, rd), img_idx=torch.zeros(1, 10), **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    assert rgb.shape == (n, 3) and bool(torch.isfinite(rgb).all())
    print(f"{n:9d} {(n + 32767) // 32768:6d} {ms:9.3f} {n / ms / 1e3:9.3f}")
    del ro, rd, rgb, disp, acc, ex
    torch.cuda.empty_cache()
