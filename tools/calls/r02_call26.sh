#!/bin/bash
# round-2 validation: smoke, full GPU test suite, the bench line (all configs) and the reference arm
mkdir -p gpurun_out
python -c "
import __graft_entry__ as g, time
t0=time.time(); g.smoke(); print('smoke s', round(time.time()-t0,2))" > gpurun_out/r02_c26_smoke.log 2>&1; tail -n 2 gpurun_out/r02_c26_smoke.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c26_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c26_gpu_tests.log
(time python bench.py) > gpurun_out/r02_c26_bench.json 2> gpurun_out/r02_c26_bench.err; tail -n 4 gpurun_out/r02_c26_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_c26_reference.json 2> gpurun_out/r02_c26_reference.err; tail -n 4 gpurun_out/r02_c26_reference.err; cut -c1-600 gpurun_out/r02_c26_reference.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c26_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['kernels'], d['roofline']['traffic'])
print('config', d['config'])
print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
for c in d['configs']:
    r=c.get('roofline') or {}
    print(c['name'], c.get('size'), c.get('route'), 'ms',c.get('ms_per_assembly'), 'frac', r.get('frac'), 'traffic', r.get('traffic'), 'first_call_s',c.get('first_call_s'), c.get('skipped'), c.get('error'))
PY
