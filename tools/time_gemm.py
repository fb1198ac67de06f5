"""Timing of nefes_linear_{fwd,dgrad,wgrad} on the SIMT fp32 GEMM and on the tf32 tensor-core GEMM at the field's layer shapes,
and of the whole field query per precision.  Usage: python tools/time_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nefes_b200 as nb
from nefes_b200 import _lib as L, ops
dev = "cuda"
lib = L.lib()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); ev0.record()
    for _ in range(n): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / n
M = 6144 * 128
for N, K in ((128, 128), (128, 191), (131, 64)):
    A = torch.randn(M, K, device=dev); Wt = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev); dA = torch.empty(M, K, device=dev); dW = torch.zeros(N, K, device=dev)
    st = L.stream_of(A)
    for mode in (0, 1):
        lib.nefes_gemm_mode(mode)
        tf = timeit(lambda: lib.nefes_linear_fwd(L.ptr(A), K, L.ptr(Wt), L.ptr(b), L.ptr(C), N, M, N, K, 1, st))
        td = timeit(lambda: lib.nefes_linear_dgrad(L.ptr(C), N, L.ptr(Wt), L.ptr(dA), K, M, N, K, None, 0, st))
        tw = timeit(lambda: lib.nefes_linear_wgrad(L.ptr(C), N, L.ptr(A), K, L.ptr(dW), M, N, K, st))
        fl = 2.0 * M * N * K
        by = 4.0 * M * (N + K)
        print(f"M={M} N={N} K={K} {'tf32' if mode else 'simt'}: fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s, {by / tf / 1e6:.0f} GB/s)  dgrad {td:.3f} ms  wgrad {tw:.3f} ms", flush=True)
lib.nefes_gemm_mode(0)
g = torch.Generator(device="cuda").manual_seed(0)
for prec, name in ((L.PREC_FP32, "fp32"), (L.PREC_TF32, "tf32"), (L.PREC_BF16, "bf16")):
    f = nb.NeRFH_NFF("fine", W=128, encode_appearance=True, encode_transient=True).cuda()
    pts = torch.rand(6144, 128, 3, device="cuda", generator=g) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(6144, 3, device="cuda", generator=g), dim=-1)
    def fwd():
        return ops.field_query(pts, dirs, f.flat, f.net_id, L.MODE_FULL, prec)
    raw = fwd(); gr = torch.randn_like(raw)
    t_f = timeit(fwd, 3)
    def both():
        f.zero_grad(); fwd().backward(gr)
    t_b = timeit(both, 3)
    print(f"field query fine 6144x128, {name}: fwd {t_f:.3f} ms, fwd+bwd {t_b:.3f} ms", flush=True)
