#!/bin/bash
# slot order modes crossed: TetP2 with coarse count classes, vector-P2 / Stokes B with the exact count order
mkdir -p gpurun_out
OUT=gpurun_out/r02_c51_sweep.jsonl
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> $OUT; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> $OUT 2>> gpurun_out/r02_c51_sweep.err; }
rm -f $OUT
run tetp2_bucket tetp2 128 SPFEM_TILE_SORT=bucket
run vecp2_exact vecp2 2048 SPFEM_TILE_SORT=1
run stokesb_exact stokesb 2048 SPFEM_TILE_SORT=1
python - <<PY
import json
for l in open('$OUT'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'), d.get('error','')[:80])
PY
tail -n 3 gpurun_out/r02_c51_sweep.err
