#!/bin/bash
# The driver's two commands at N=1, as it runs them at round end.
mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_driverlike_ref.json 2> gpurun_out/r2_driverlike_ref.err ) 2> gpurun_out/r2_driverlike_ref.time
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_driverlike.json 2> gpurun_out/r2_driverlike.err ) 2> gpurun_out/r2_driverlike.time
cat gpurun_out/r2_driverlike_ref.time gpurun_out/r2_driverlike.time; tail -c 600 gpurun_out/r2_driverlike.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_driverlike.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_driverlike_ref.json').read().strip().splitlines()[-1])
print('headline', d['metric'], round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'of', round(d['roofline']['peak'],2), 'burst', round(d['roofline']['peak_burst'],2), 'e2e', round(d['e2e']['value'],2), 'ref', round(r['value'],3), 'same_config', d['config']==r['config'], 'total', d['total_run_s'])
for k,v in d['configs'].items():
    rf=v['roofline']; e=v.get('e2e') or {}; c=v.get('cpu_baseline') or {}
    print(k, round(v['value'],2), v['unit'], 'ms', round(v['ms_per_step'],3), 'frac', round(rf['frac'],3), 'e2e', round(e.get('value',0),2), 'cpu', round(c.get('value',0),3), 'ref-arm', round(r['configs'][k]['value'],3), v['clocks'].get('reasons'), v.get('extra'))
PY
