#!/bin/bash
# A/B of an environment switch of the library: `tools/gpu_ab_env.sh VAR a b` runs the short bench with VAR=a and VAR=b,
# interleaved, three repetitions, and prints the step and the per-kernel times
VAR=$1; A=$2; B=$3
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in $A $B; do
  env $VAR=$v timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 8 > gpurun_out/ab_env.json 2>/dev/null
  echo "$VAR=$v (rep $rep): $(python tools/bench_line.py gpurun_out/ab_env.json 2>/dev/null | sed -n 1,1p | cut -c1-44) $(python tools/bench_line.py gpurun_out/ab_env.json 2>/dev/null | sed -n 2,9p | tr ' ' '\n' | grep 'gemm_\|attention' | tr '\n' ' ')"
done
done
