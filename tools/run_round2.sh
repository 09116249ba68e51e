#!/bin/bash
# Round-2 GPU pass: parity tests, smoke, the default bench line (all five configurations) and the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -25 gpurun_out/r2_tests_gpu.log; cat gpurun_out/r2_smoke.log; tail -c 1500 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
    print('total_run_s', d.get('total_run_s'))
    for k,v in d['configs'].items():
        r=v['roofline']; e=v.get('e2e') or {}; c=v.get('cpu_baseline') or {}
        print(k, round(v['value'],2), v['unit'], 'ms', round(v['ms_per_step'],3), 'frac', round(r['frac'],3), r['bound'], 'e2e', round(e.get('value',0),2), 'cpu', round(c.get('value',0),3), v['clocks'].get('reasons'), v.get('verified'))
except Exception as ex:
    print('ERR', ex)
PY
