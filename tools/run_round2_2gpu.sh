#!/bin/bash
# Two-GPU pass: the new parity tests (incl. the CUDA+NCCL m-sharded sketch under torchrun and the single-process
# rb_comm_init form in the C++ drop-in test) and the bench at N=2.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_cpp_dropin.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -40 gpurun_out/r2_tests_2gpu.log; tail -c 1500 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print('total_run_s', d.get('total_run_s'), d.get('comm'), d.get('numa_cpus_rank0'))
    for k,v in d['configs'].items():
        r=v['roofline']; e=v.get('e2e') or {}; c=v.get('cpu_baseline') or {}
        print(k, round(v['value'],2), v['unit'], 'ms', round(v['ms_per_step'],3), 'frac', round(r['frac'],3), r['bound'], 'e2e', round(e.get('value',0),2), 'cpu', round(c.get('value',0),3), v['clocks'].get('reasons'), v.get('verified'))
except Exception as ex:
    print('ERR', ex)
PY
