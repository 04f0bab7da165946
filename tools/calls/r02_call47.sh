#!/bin/bash
# round-end evidence on the final code: bench line, full GPU test suite, smoke, launch list of the
# bench command, full captures of the element-tile kernel for Q2 (parallelogram branch, row records)
# and vector-P2 (count-ordered slots, unchanged format)
mkdir -p gpurun_out
(time python bench.py) > gpurun_out/r02_c47_bench.json 2> gpurun_out/r02_c47_bench.err; tail -n 4 gpurun_out/r02_c47_bench.err
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/r02_c47_gpu_tests.log 2>&1; tail -n 2 gpurun_out/r02_c47_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c47_smoke.log 2>&1; tail -n 2 gpurun_out/r02_c47_smoke.log
python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/r02_c47_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final2_launches.csv \
    python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/r02_c47_ncu_bench.log 2>&1
prof () { name=$1; kind=$2; size=$3; pat=$4; n=$5
  python tools/scale_probe.py $kind $size > gpurun_out/r02_c47_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c $n -f -o gpurun_out/r02_final2_$name \
      python tools/scale_probe.py $kind $size > gpurun_out/r02_c47_ncu_$name.log 2>&1
  cut -c1-200 gpurun_out/r02_c47_plain_$name.log; tail -n 2 gpurun_out/r02_c47_ncu_$name.log; }
prof q2_tile q2 10 spfem_elem_tile 1
ls -la gpurun_out/r02_final2*.ncu-rep
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c47_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['kernels'], d['roofline']['traffic'])
print('e2e', d['e2e'])
for c in d['configs']:
    r=c.get('roofline') or {}
    print(c['name'], c.get('size'), c.get('route'), 'ms',c.get('ms_per_assembly'), 'frac', r.get('frac'), 'first_call_s',c.get('first_call_s'), c.get('skipped'), c.get('error'))
PY
