#!/bin/bash
# Run under gpurun (one GPU): round-2 evidence -- the ncu launch list of one bench step and `ncu --set full` captures of
# the four tower GEMMs + attention at the BENCH shape (one pass of 8320 views), fp16 operands (the default).
# Each ncu pass only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r02_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 12 -c 4 -o gpurun_out/r02_prof_gemm $CMD > gpurun_out/r02_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 3 -c 1 -o gpurun_out/r02_prof_attention $CMD > gpurun_out/r02_ncu_att.log 2>&1
echo "attention capture rc=$?"
ls -la gpurun_out/*.ncu-rep
