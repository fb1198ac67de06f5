"""Pure-write, pure-read and copy HBM bandwidth (torch kernels, CUDA events, best of 5) -- roofline denominators."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def best(f, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
a.zero_(); b.zero_(); torch.cuda.synchronize()
t = best(lambda: a.zero_()); print(f"write (zero_ 4 GiB): {4 * n / t / 1e6:.0f} GB/s")
t = best(lambda: a.sum()); print(f"read (sum 4 GiB): {4 * n / t / 1e6:.0f} GB/s")
t = best(lambda: b.copy_(a)); print(f"copy (4 GiB -> 4 GiB): {8 * n / t / 1e6:.0f} GB/s read+write")
