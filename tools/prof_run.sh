#!/bin/bash
# ncu --set full captures of single kernels (args: list of prof_kernels.py targets). Reports land in gpurun_out/.
mkdir -p gpurun_out
for t in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_$t \
    --launch-skip 0 -c 12 python tools/prof_kernels.py $t 1 > gpurun_out/prof_$t.log 2>&1
  tail -2 gpurun_out/prof_$t.log
done
