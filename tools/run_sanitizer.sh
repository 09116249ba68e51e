#!/bin/bash
# compute-sanitizer (memcheck) over the kernels added or changed in round 2, at small shapes.
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 \
    python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x \
    -k "random_coo or column_block or short_axis or one_rank or conversions_vs_reference or tensor_core_float_sketch_vs_oracle or cluster_and_halves or dmma_double or materialised" \
    > gpurun_out/r2_sanitizer.log 2>&1
echo "exit code $?" >> gpurun_out/r2_sanitizer.log
grep -E "ERROR SUMMARY|passed|failed|exit code|Invalid|Error" gpurun_out/r2_sanitizer.log | head -20
