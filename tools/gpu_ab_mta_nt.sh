#!/bin/bash
# Run under gpurun: thread count of the small-V MTA kernel (JCB_MTA_NT) at N = 1 and N = 16 crops, device-resident step only
set -u
mkdir -p gpurun_out
for cfg in 1:4160 16:489; do
  set -- ${cfg%%:*} ${cfg##*:}
  for nt in 128 64 32; do
    JCB_MTA_NT=$nt timeout 300 python bench.py --crops $1 --images-per-gpu $2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/mta_nt.json 2>/dev/null
    python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/mta_nt.json") if l.startswith("{")][-1])
print("N=$1 crops NT=$nt: %.0f images/s, %.2f ms/step, mta %.3f ms" % (d["value"], d["ms_per_step"], d["roofline"]["per_kernel"]["mta"]["ms_per_step"]))
PY
  done
done
python -m pytest tests/test_gpu_mta_head.py -q 2>&1 | tail -2
JCB_MTA_NT=32 python -m pytest tests/test_gpu_mta_head.py -q -k "mta" 2>&1 | tail -2
JCB_MTA_NT=64 python -m pytest tests/test_gpu_mta_head.py -q -k "mta" 2>&1 | tail -2
