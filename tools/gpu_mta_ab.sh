#!/bin/bash
# Run under gpurun: the MTA share of a pipeline call, this build against jittor-clip-fewshot_b200/csrc/build/libjclip_old.so
set -u
OLD=$PWD/jittor-clip-fewshot_b200/csrc/build/libjclip_old.so
for rep in 1 2; do
  for cfg in "128 65" "489 17" "4160 2" "1 65" "3 65"; do
    echo "new: $(python tools/mta_probe.py $cfg)"
    [ -f "$OLD" ] && echo "old: $(JCB_LIB_PATH=$OLD python tools/mta_probe.py $cfg)"
  done
done
