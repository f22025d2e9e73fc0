#!/bin/bash
# Run under gpurun (one GPU): full GPU parity suite, smoke(), the default bench line, the reference arm,
# then (each only after its plain command exited 0) the ncu launch list of the bench command and one
# --set full capture of the MTA solver.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_default.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD="python tools/mta_probe.py 128 65"
$CMD > gpurun_out/mta_probe_plain.log 2>&1 && cat gpurun_out/mta_probe_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:mta_fast_kernel -s 2 -c 1 -o gpurun_out/mta_fast_prof $CMD > gpurun_out/mta_fast_ncu.log 2>&1
echo "mta capture rc=$?"
