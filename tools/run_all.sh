#!/bin/bash
# Full GPU pass: parity tests, smoke, and one bench line per workload (results land in gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for w in c2 c1 c3 c4 c5; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
tail -5 gpurun_out/tests_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench_c*.json
