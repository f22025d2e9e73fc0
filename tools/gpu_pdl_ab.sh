#!/bin/bash
# Run under gpurun (one GPU): programmatic dependent launch on / off (JCB_PDL): the whole GPU suite with it on, then
# the single-image call and a short device-resident bench both ways, interleaved
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for rep in 1 2; do
  for pdl in 0 1; do
    echo "== JCB_PDL=$pdl (rep $rep)"
    JCB_PDL=$pdl python tools/single_image_latency.py
    JCB_PDL=$pdl python bench.py --no-cpu-baseline --no-e2e --steps 6 > gpurun_out/ab_pdl$pdl.json 2>/dev/null; python tools/bench_line.py gpurun_out/ab_pdl$pdl.json 2>/dev/null | sed -n 1,1p | cut -c1-120
  done
done
