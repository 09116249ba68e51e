#!/bin/bash
# Full GPU pass: parity tests, smoke, and one bench line per workload (results land in gpurun_out/).
# usage: tools/run_all.sh [workloads...]   (default: c2 c1 c3 c4 c5)
mkdir -p gpurun_out
WL=${@:-c2 c1 c3 c4 c5}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for w in $WL; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
tail -3 gpurun_out/tests_gpu.log; cat gpurun_out/smoke.log
for w in $WL; do tail -c 600 gpurun_out/bench_$w.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
    r=d['roofline']; e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
    print('$w', round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), r['bound'], 'e2e', e.get('value'), 'cpu', c.get('value'), d['clocks'].get('reasons'))
except Exception as ex:
    print('$w', 'ERR', ex)
PY
done
