#!/usr/bin/env python
"""Developer tool: read-only, write-only and copy HBM bandwidth of this GPU with plain torch ops (CUDA events, best of
10), to put the write-dominated kernels (length regulator) against the right roofline."""
import torch

n = 1 << 30  # 1 Gi bf16 elements = 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
a.fill_(1.0)
torch.cuda.synchronize()


def best(fn, bytes_moved, reps=10):
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return bytes_moved / (min(t) * 1e-3) / 1e9


print(f"copy  (read+write) : {best(lambda: b.copy_(a), 4 * n):8.1f} GB/s")
print(f"write (fill_)      : {best(lambda: b.fill_(2.0), 2 * n):8.1f} GB/s")
print(f"write (zero_)      : {best(lambda: b.zero_(), 2 * n):8.1f} GB/s")
ai = a.view(torch.int32)
print(f"read  (sum int32)  : {best(lambda: ai.sum(), 2 * n):8.1f} GB/s")
