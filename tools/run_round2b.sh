#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "short_axis or example" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2b_tests.log
tail -40 gpurun_out/r2b_tests.log
python tools/exp_dense_layouts.py > gpurun_out/r2b_layouts.log 2>&1; tail -20 gpurun_out/r2b_layouts.log
NCU_FILTER="-k regex:spdata" bash tools/run_evidence_one.sh sksp
