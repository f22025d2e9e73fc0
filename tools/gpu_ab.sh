#!/bin/bash
# A/B of library variants (built with `build.py --variant`): GEMM micro-benchmarks and the full bench, interleaved
V=jittor-clip-fewshot_b200/csrc/build/variants
for rep in 1 2; do
for lib in default $@; do
  if [ $lib = default ]; then unset JCB_LIB_PATH; else export JCB_LIB_PATH=$PWD/$V/$lib.so; fi
  echo "== $lib (rep $rep)"
  if [ -z "$NO_MICRO" ]; then timeout 200 python tools/bench_kernel.py gemm 8320 2>&1 | tail -4; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 6 > gpurun_out/ab_$lib.json 2>/dev/null; python tools/bench_line.py gpurun_out/ab_$lib.json | sed -n 1,2p
done
done
