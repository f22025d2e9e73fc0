#!/bin/bash
# Run under gpurun: ncu --set full of the non-GEMM kernels at the BENCH shape + a launch list of one step
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_small.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-attention_tcgen05|mta_kernel|mta_probs|head_kernel}" -s ${SKIP:-11} -c ${COUNT:-4} -o gpurun_out/prof_${TAG:-small} $CMD > gpurun_out/ncu_small.log 2>&1
echo "full capture rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-70} -c 80 --csv --log-file gpurun_out/launches_${TAG:-small}.csv $CMD > gpurun_out/ncu_launch_small.log 2>&1
echo "launch list rc=$?"
