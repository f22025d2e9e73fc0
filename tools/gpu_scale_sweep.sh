#!/bin/bash
# Run under gpurun --gpus 8: the 1 -> 8 GPU curve of the default bench on ONE box, back to back
set -u
mkdir -p gpurun_out
PORT=29560
for G in 1 2 4 8; do
  out=gpurun_out/scale_n$G.json
  if [ "$G" = 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/scale_err.log
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/scale_err.log
  fi
  PORT=$((PORT + 1))
  python - <<PY
import json
d = json.loads([l for l in open("$out") if l.startswith("{")][-1])
print("N=%d: %.0f images/s (%.2f ms/step), e2e %.0f, from images %.0f, shard_check %s" % (
    d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_from_images"]["value"],
    (d["config"].get("shard_check") or {}).get("bit_identical_by_rank")))
PY
done
