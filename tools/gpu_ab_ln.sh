#!/bin/bash
# A/B of out_proj (EPI_RESID_LNPREP_SHORT) staging variants: micro-benchmark + short bench, interleaved, two repetitions
V=jittor-clip-fewshot_b200/csrc/build/variants
mkdir -p gpurun_out
for rep in 1 2; do
for lib in default "$@"; do
  if [ $lib = default ]; then unset JCB_LIB_PATH; else export JCB_LIB_PATH=$PWD/$V/$lib.so; fi
  echo "== $lib (rep $rep)"
  timeout 200 python tools/bench_kernel.py gemm_ln 8320 2>&1 | tail -2
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 6 > gpurun_out/ab_$lib.json 2>/dev/null; python tools/bench_line.py gpurun_out/ab_$lib.json | sed -n 1,1p; python tools/bench_line.py gpurun_out/ab_$lib.json | grep gemm_out | tr ' ' '\n' | grep "gemm_out\|attention"
done
done
