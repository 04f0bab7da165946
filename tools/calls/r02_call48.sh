#!/bin/bash
# N GPUs (N = number visible): NCCL interface-exchange tests + the multi-GPU bench line
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_nccl_exchange.py -m gpu -x -q > gpurun_out/r02_c48_nccl_tests_${N}gpu.log 2>&1
tail -n 5 gpurun_out/r02_c48_nccl_tests_${N}gpu.log
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 \
    bench.py --gpus $N --steps 100 --warmup 3) > gpurun_out/r02_c48_bench_${N}gpu.json 2> gpurun_out/r02_c48_bench_${N}gpu.err
tail -n 6 gpurun_out/r02_c48_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_c48_bench_${N}gpu.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','parity_checked','exchange')})
    print(d['e2e'])
    for c in d['configs']:
        r=c.get('roofline') or {}
        print(c['name'], c.get('size'), c.get('route'), 'ms',c.get('ms_per_assembly'), 'frac', r.get('frac'), 'n_per_gpu', c.get('n_per_gpu'), 'first_call_s',c.get('first_call_s'), 'mesh_s',c.get('mesh_s'), c.get('skipped'), c.get('error'))
except Exception as e:
    print('no line', e)
PY
