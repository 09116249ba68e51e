#!/bin/bash
# one ncu --set full capture: tools/run_evidence_one.sh <prof_kernels target>   (NCU_FILTER="-k regex:name" optional)
t=$1
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_prof_$t ${NCU_FILTER:-} \
    --launch-skip 0 -c 4 python tools/prof_kernels.py $t 1 > gpurun_out/r2_prof_$t.log 2>&1
tail -2 gpurun_out/r2_prof_$t.log
