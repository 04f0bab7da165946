#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c35_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c35_gpu_tests.log
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c35_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c35_sweep.jsonl 2>> gpurun_out/r02_c35_sweep.err; }
rm -f gpurun_out/r02_c35_sweep.jsonl
run stokesb stokesb 3162 X=1
run trip2 trip2 3162 X=1
run vec vecp2 3162 X=1
run q1 q1 11 X=1
python - <<PY
import json
for l in open('gpurun_out/r02_c35_sweep.jsonl'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'),'prep',d.get('prepare_s'), d.get('error','')[:80])
PY
