#!/bin/bash
# Run under gpurun: BASELINE config 5's crop sweep, N in {1, 16, 64} crops per image at ~8.3 k views per GPU per step.
#   GPUS=8 tools/gpu_sweep_crops.sh   (default 1)
set -u
mkdir -p gpurun_out
G=${GPUS:-1}
PORT=29540
for cfg in ${CONFIGS:-1:4160 16:489 64:128}; do      # crops:images-per-GPU-per-step
  set -- ${cfg%%:*} ${cfg##*:}
  out=gpurun_out/sweep_g${G}_n$1.json
  if [ "$G" = 1 ]; then
    timeout 600 python bench.py --crops $1 --images-per-gpu $2 --steps 10 --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/sweep_err.log
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $G --crops $1 --images-per-gpu $2 --steps 10 --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/sweep_err.log
  fi
  PORT=$((PORT + 1))
  python - <<PY
import json
d = json.loads([l for l in open("$out") if l.startswith("{")][-1])
print("N=$1 crops, $2 images/GPU/step, %d GPU: %.0f images/s (%.0f views/s), e2e %.0f, from images %.0f, %.2f ms/step, step frac %.3f" % (
    d["n_gpus"], d["value"], d["value"] * ($1 + 1), d["e2e"]["value"], (d.get("e2e_from_images") or {}).get("value", 0), d["ms_per_step"],
    d["roofline"]["whole_step_frac"]))
PY
done
