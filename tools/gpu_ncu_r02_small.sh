#!/bin/bash
# Run under gpurun (one GPU): ncu --set full of the kernels that are new in round 2 (tools/r02_small_workload.py), one or two
# launches of each
set -u
mkdir -p gpurun_out
CMD="python tools/r02_small_workload.py"
$CMD > gpurun_out/r02_small_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_small_plain.log; exit 1; }
i=0
for k in attention_tcgen05 resample_h_kernel resample_v_kernel mta_split_x mta_softmax_rows mta_fast_kernel; do
  i=$((i + 1))
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r02_prof_small_$i $CMD > gpurun_out/r02_ncu_small_$i.log 2>&1
  echo "$k capture rc=$?"
done
