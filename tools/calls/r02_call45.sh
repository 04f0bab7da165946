#!/bin/bash
# row records + chunk headers instead of per-slot positions (slots in CSR order):
# 2-D scalar plans use it by default; vector plans / TetP2 with SPFEM_TILE_SORT=0 against their sorted default
mkdir -p gpurun_out
OUT=gpurun_out/r02_c45_sweep.jsonl
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> $OUT; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> $OUT 2>> gpurun_out/r02_c45_sweep.err; }
rm -f $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile or fused or mode or dense or packed or scalar_slots" > gpurun_out/r02_c45_gpu_tests.log 2>&1; tail -n 2 gpurun_out/r02_c45_gpu_tests.log
run q2 q2 11 X=1
run trip2 trip2 2048 X=1
run vecp2 vecp2 2048 X=1
run vecp2_csr vecp2 2048 SPFEM_TILE_SORT=0
run stokesb stokesb 2048 X=1
run stokesb_csr stokesb 2048 SPFEM_TILE_SORT=0
run tetp2 tetp2 128 X=1
run tetp2_csr tetp2 128 SPFEM_TILE_SORT=0
python - <<PY
import json
for l in open('$OUT'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'), 'blob/elt', round(t.get('blob_bytes',0)/max(d.get('nt',1),1),1), d.get('error','')[:80])
PY
tail -n 3 gpurun_out/r02_c45_sweep.err
