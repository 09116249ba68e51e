#!/bin/bash
# N-GPU bench (N = number of visible GPUs): the default line (all five configurations) and the reference arm.
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1200 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N total_run_s', d.get('total_run_s'), d.get('comm'), d.get('numa_cpus_rank0'))
    for k,v in d['configs'].items():
        r=v['roofline']; e=v.get('e2e') or {}; c=v.get('cpu_baseline') or {}
        print(k, round(v['value'],2), v['unit'], 'ms', round(v['ms_per_step'],3), 'frac', round(r['frac'],3), r['bound'], 'e2e', round(e.get('value',0),2), 'pcie/rank', round(e.get('pcie_gbs_per_rank',0),1), 'cpu', round(c.get('value',0),3), v['clocks'].get('reasons'), v.get('verified'))
except Exception as ex:
    print('ERR', ex)
PY
