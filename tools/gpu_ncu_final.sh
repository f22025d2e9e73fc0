#!/bin/bash
# Run under gpurun (one GPU): `ncu --set full` of the final build -- the four tower GEMMs at the BENCH shape (the kernel
# gained a second operand pair and programmatic dependent launch since the r02i capture), and the cluster forms of the
# tail / head inside a one-image call.  Each ncu pass only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/final_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/final_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 12 -c 4 -o gpurun_out/final_prof_gemm $CMD > gpurun_out/final_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
CMD2="python tools/single_image_profile.py 65"
$CMD2 > gpurun_out/final_single_plain.log 2>&1 || { echo "plain single-image run failed"; tail -5 gpurun_out/final_single_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'tail_kernel|head_kernel' -s 4 -c 2 -o gpurun_out/final_prof_small $CMD2 > gpurun_out/final_ncu_small.log 2>&1
echo "small capture rc=$?"
ls -la gpurun_out/*.ncu-rep
