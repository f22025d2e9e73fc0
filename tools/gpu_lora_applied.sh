#!/bin/bash
# Run under gpurun (one GPU): the applied-LoRA tests first, then the whole GPU suite (the GEMM kernel gained a second
# operand pair), then the default bench (its lora_applied leg + a check that the merged-mode GEMM rates did not move)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lora_applied.py -x -q > gpurun_out/pytest_lora_applied.log 2>&1; echo "applied rc=$?"; tail -30 gpurun_out/pytest_lora_applied.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
if [ -z "${SKIP_BENCH:-}" ]; then python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; fi
python tools/bench_line.py gpurun_out/bench_default.json 2>/dev/null | head -5
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['lora_applied'], d['other_operand_type']['value'])
r = d['roofline']; print(r['achieved'], r['library_same_shape']['tflops'], {k: round(v['tflops'], 1) for k, v in r['per_kernel'].items() if 'tflops' in v})
PY
