#!/bin/bash
# ncu evidence of the round's final kernels: launch list of the bench command, full captures of
# the vertex-fan kernel (config 2) and the element-tile kernel (vector-P2, TetP2, TriP2)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/r02_c39_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/r02_c39_ncu_bench.log 2>&1
prof () { name=$1; kind=$2; size=$3; pat=$4; n=$5
  python tools/scale_probe.py $kind $size > gpurun_out/r02_c39_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c $n -f -o gpurun_out/r02_final_$name \
      python tools/scale_probe.py $kind $size > gpurun_out/r02_c39_ncu_$name.log 2>&1
  cut -c1-200 gpurun_out/r02_c39_plain_$name.log; tail -n 2 gpurun_out/r02_c39_ncu_$name.log; }
prof fan trip1 11 spfem_elem_fan 2
prof vecp2_tile vecp2 2048 spfem_elem_tile 1
prof tetp2_tile tetp2 128 spfem_elem_tile 1
prof trip2_tile trip2 2048 spfem_elem_tile 1
ls -la gpurun_out/*.ncu-rep
(time python bench.py) > gpurun_out/r02_c39_bench.json 2> gpurun_out/r02_c39_bench.err; tail -n 4 gpurun_out/r02_c39_bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_c39_reference.json 2> gpurun_out/r02_c39_reference.err; tail -n 2 gpurun_out/r02_c39_reference.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c39_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')})
print('roofline', d['roofline']['frac'], d['roofline']['kernels'], d['roofline']['traffic'])
print('e2e', d['e2e'])
for c in d['configs']:
    r=c.get('roofline') or {}
    print(c['name'], c.get('size'), c.get('route'), 'ms',c.get('ms_per_assembly'), 'frac', r.get('frac'), 'first_call_s',c.get('first_call_s'), c.get('skipped'), c.get('error'))
PY
