#!/bin/bash
# second batch of ncu --set full captures (gpurun copies back at most 64 MiB per call)
mkdir -p gpurun_out
for t in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_prof_$t ${NCU_FILTER:-} \
    --launch-skip 0 -c 8 python tools/prof_kernels.py $t 1 > gpurun_out/r2_prof_$t.log 2>&1
  tail -1 gpurun_out/r2_prof_$t.log
done
NCU_FILTER="-k regex:spdata" bash tools/run_evidence_one.sh sksp
ls -la gpurun_out/*.ncu-rep
