#!/bin/bash
# Evidence pass (one GPU): the whole GPU suite, the ncu launch list of the default bench command, and ncu --set full
# captures of the current default kernels. Everything lands in gpurun_out/ (copied to profiles/ from there).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_tests_gpu_full.log
tail -5 gpurun_out/r2_tests_gpu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/r2_bench_under_ncu.log
for t in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_prof_$t \
    ${NCU_FILTER:-} --launch-skip 0 -c 8 python tools/prof_kernels.py $t 1 > gpurun_out/r2_prof_$t.log 2>&1
  tail -2 gpurun_out/r2_prof_$t.log
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
