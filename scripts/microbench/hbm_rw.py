#!/usr/bin/env python
"""HBM bandwidth by access mix (torch kernels, CUDA events): write-only (memset / fill), read-only (sum), copy."""
import torch

n = 1 << 30                      # 1 Gi floats = 4 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


gb = n * 4 / 1e9
print(f"write-only  zero_   : {gb / timeit(lambda: a.zero_()) :.1f} GB/ms -> x1000 GB/s")
print(f"write-only  fill_   : {gb / timeit(lambda: a.fill_(1.5)):.1f}")
print(f"read-only   sum     : {gb / timeit(lambda: a.sum()):.1f}")
print(f"copy (r+w)  copy_   : {2 * gb / timeit(lambda: b.copy_(a)):.1f}")
print(f"r+w in place mul_   : {2 * gb / timeit(lambda: a.mul_(1.0001)):.1f}")
