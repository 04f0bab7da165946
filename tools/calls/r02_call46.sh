#!/bin/bash
# Q2 / Q1 after the parallelogram branch: threads per element and tile size
mkdir -p gpurun_out
OUT=gpurun_out/r02_c46_sweep.jsonl
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> $OUT; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> $OUT 2>> gpurun_out/r02_c46_sweep.err; }
rm -f $OUT
run q2_np1 q2 11 SPFEM_TILE_NPART=1
run q2_np1_e128 q2 11 SPFEM_TILE_NPART=1 SPFEM_TILE_E=128
run q2_np1_e64 q2 11 SPFEM_TILE_NPART=1 SPFEM_TILE_E=64
run q2_e128 q2 11 SPFEM_TILE_E=128
run q2_e64 q2 11 SPFEM_TILE_E=64
run q1_e128 q1 11 SPFEM_TILE_E=128
run q1_e384 q1 11 SPFEM_TILE_E=384
python - <<PY
import json
for l in open('$OUT'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'), 'blob/elt', round(t.get('blob_bytes',0)/max(d.get('nt',1),1),1), d.get('error','')[:80])
PY
tail -n 3 gpurun_out/r02_c46_sweep.err
