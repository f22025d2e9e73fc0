#!/usr/bin/env python
"""Context for roofline.frac: what cuBLAS (torch.matmul, bf16 in / fp32 accumulate / bf16 out) sustains on THIS path's GEMM
shapes, back to back for ~2 s each under the same power cap -- next to the 8192^3 figure MEASURED_PEAKS.json quotes.
A plain library GEMM: no bias, activation, residual or LayerNorm work in its epilogue.
    python tools/cublas_shape_probe.py [n_views]"""
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8320
M = n * 50
dev = torch.device("cuda", 0)
shapes = {"8192^3": (8192, 8192, 8192), "qkv   M x 2304 x 768": (M, 2304, 768), "out   M x 768 x 768": (M, 768, 768),
          "c_fc  M x 3072 x 768": (M, 3072, 768), "c_proj M x 768 x 3072": (M, 768, 3072)}
for name, (m, nn, k) in shapes.items():
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = torch.randn(nn, k, device=dev).to(torch.bfloat16)
    out = torch.empty(m, nn, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize()
    iters = max(int(2.0 / (2.0 * m * nn * k / 1.2e15)), 5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"cuBLAS bf16 {name:24s} {ms:8.3f} ms  {2.0 * m * nn * k / ms / 1e9:7.0f} TFLOP/s  ({iters} launches back to back)")
    del a, w, out
