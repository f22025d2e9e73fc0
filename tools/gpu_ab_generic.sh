#!/bin/bash
# A/B of library variants (build.py --variant): GEMM micro-benchmarks and a short bench, interleaved, two repetitions
V=jittor-clip-fewshot_b200/csrc/build/variants
mkdir -p gpurun_out
for rep in 1 2; do
for lib in default "$@"; do
  if [ $lib = default ]; then unset JCB_LIB_PATH; else export JCB_LIB_PATH=$PWD/$V/$lib.so; fi
  echo "== $lib (rep $rep)"
  timeout 200 python tools/bench_kernel.py gemm 8320 2>&1 | grep fc1
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 6 > gpurun_out/ab_$lib.json 2>/dev/null; python tools/bench_line.py gpurun_out/ab_$lib.json 2>/dev/null | sed -n 1,1p | cut -c1-100; python tools/bench_line.py gpurun_out/ab_$lib.json 2>/dev/null | tr ' ' '\n' | grep "gemm_fc1"
done
done
