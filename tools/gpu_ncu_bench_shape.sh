#!/bin/bash
# Run under gpurun: ncu --set full of the GEMM kernels at the BENCH shape (one pass of 8320 views), for roofline.traffic
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_bs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-gemm_tcgen05_2cta} -s ${SKIP:-12} -c ${COUNT:-4} -o gpurun_out/prof_${TAG:-gemm_benchshape} $CMD > gpurun_out/ncu_bs.log 2>&1
echo "full capture rc=$?"
