#!/bin/bash
# Run under gpurun: chunk-size sweep, ncu launch list, one full ncu capture of the GEMM kernel.
set -u
mkdir -p gpurun_out
for c in ${CHUNKS:-2048 4160 8320}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --chunk-views $c --no-cpu-baseline --no-e2e 2>/dev/null > gpurun_out/chunk_$c.json
  python - <<PY
import json
d=json.load(open("gpurun_out/chunk_$c.json"))
pk=d["roofline"]["per_kernel"]
print("chunk $c: %.1f img/s, %.2f ms/step, gemm %.0f TF/s | "%(d["value"],d["ms_per_step"],d["roofline"]["achieved"]) + " ".join("%s=%.1f"%(k,v["ms_per_step"]) for k,v in pk.items()))
PY
done
CMD="python bench.py --steps 2 --warmup 1 --images-per-gpu 32 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-gemm_bf16} -s ${SKIP:-40} -c ${COUNT:-4} -o gpurun_out/prof_${TAG:-gemm} $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
