#!/bin/bash
# Run under gpurun: MTA parity tests + the MTA share of a pipeline call at the headline shape and the small-V shapes
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mta_head.py -q -x 2>&1 | tail -3
python tools/mta_probe.py 128 65
python tools/mta_probe.py 489 17
python tools/mta_probe.py 4160 2
python tools/mta_probe.py 1 65
python tools/bench_kernel.py mta 128
