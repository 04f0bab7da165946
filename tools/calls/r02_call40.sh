#!/bin/bash
# block plans: one 8-byte record per slot
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c40_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c40_sweep.jsonl 2>> gpurun_out/r02_c40_sweep.err; }
rm -f gpurun_out/r02_c40_sweep.jsonl
run vec vecp2 2048 X=1
run stokesb stokesb 2048 X=1
run vec_full vecp2 3162 X=1
run stokesb_full stokesb 3162 X=1
python - <<PY
import json
for l in open('gpurun_out/r02_c40_sweep.jsonl'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'),'blob',t.get('blob_bytes'), d.get('error','')[:80])
PY
tail -n 3 gpurun_out/r02_c40_sweep.err
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c40_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c40_gpu_tests.log
