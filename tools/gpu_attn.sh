#!/bin/bash
# attention kernel: parity tests, then micro-benchmarks of both implementations
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -15 | tee gpurun_out/attn_test.log
timeout 300 python tools/bench_kernel.py attention 8320 2>&1 | tail -2 | tee gpurun_out/attn_bench_tc.log
JCB_ATT_IMPL=mma timeout 300 python tools/bench_kernel.py attention 8320 2>&1 | tail -2 | tee gpurun_out/attn_bench_mma.log
