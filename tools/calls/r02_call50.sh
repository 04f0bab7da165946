#!/bin/bash
# final code: bench line, full GPU suite, smoke, launch list of the bench command INCLUDING the timed
# region (set-up launches skipped)
mkdir -p gpurun_out
(time python bench.py) > gpurun_out/r02_c50_bench.json 2> gpurun_out/r02_c50_bench.err; tail -n 4 gpurun_out/r02_c50_bench.err
(time timeout 900 python -m pytest tests/ -m gpu -x -q) > gpurun_out/r02_c50_gpu_tests.log 2>&1; tail -n 5 gpurun_out/r02_c50_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c50_smoke.log 2>&1; tail -n 2 gpurun_out/r02_c50_smoke.log
CMD="python bench.py --steps 5 --warmup 3 --configs none --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/r02_c50_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_final3_launches.csv \
    $CMD > gpurun_out/r02_c50_ncu_bench.log 2>&1
tail -n 1 gpurun_out/r02_c50_ncu_bench.log | cut -c1-300
grep -c spfem_elem_fan gpurun_out/r02_final3_launches.csv
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c50_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['kernels'], d['roofline']['traffic'])
print('e2e', d['e2e'])
for c in d['configs']:
    r=c.get('roofline') or {}
    print(c['name'], c.get('size'), c.get('route'), 'ms',c.get('ms_per_assembly'), 'frac', r.get('frac'), 'first_call_s',c.get('first_call_s'), c.get('skipped'), c.get('error'))
PY
