#!/usr/bin/env python
"""Host->device copy bandwidth from pinned memory on this box: idle GPU, and while a GEMM loop keeps the GPU busy."""
import time
import torch
dev = torch.device("cuda", 0)
n = 1252392960
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(3)
d = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
def copy_ms(reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
    s.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
copy_ms(1)
ms = copy_ms()
print(f"idle: {ms:.1f} ms  {n / ms / 1e6:.1f} GB/s")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(200):
    c = a @ b
ms = copy_ms()
torch.cuda.synchronize()
print(f"under GEMM load: {ms:.1f} ms  {n / ms / 1e6:.1f} GB/s")
for sz in (64 << 20, 256 << 20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for off in range(0, n - sz, sz):
            d[off:off + sz].copy_(h[off:off + sz], non_blocking=True)
    s.synchronize()
    dt = time.perf_counter() - t0
    print(f"chunks of {sz >> 20} MiB: {(n // sz) * sz / dt / 1e9:.1f} GB/s")
