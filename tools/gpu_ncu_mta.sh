#!/bin/bash
# Run under gpurun (one GPU): launch list + one ncu --set full capture (source lines) of the MTA solver inside a pipeline call
set -u
mkdir -p gpurun_out
CMD="python tools/mta_probe.py 128 65"
$CMD > gpurun_out/mta_probe_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/mta_probe_plain.log; exit 1; }
cat gpurun_out/mta_probe_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mta -c 40 --csv --log-file gpurun_out/mta_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mta_fast_kernel -s 2 -c 1 -o gpurun_out/mta_fast_prof $CMD > gpurun_out/mta_fast_ncu.log 2>&1
echo "capture rc=$?"
