#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c32_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c32_sweep.jsonl 2>> gpurun_out/r02_c32_sweep.err; }
rm -f gpurun_out/r02_c32_sweep.jsonl
run base vecp2 2048 X=1
run e48 vecp2 2048 SPFEM_TILE_E=48
run e56 vecp2 2048 SPFEM_TILE_E=56
run e72 vecp2 2048 SPFEM_TILE_E=72
run e80 vecp2 2048 SPFEM_TILE_E=80
run t192 vecp2 2048 SPFEM_TILE_THREADS=192
run t320 vecp2 2048 SPFEM_TILE_T=320 SPFEM_TILE_THREADS=320 SPFEM_TILE_MINB=2
run t128 vecp2 2048 SPFEM_TILE_THREADS=128
run sort0 vecp2 2048 SPFEM_TILE_SORT=0
run sort1 vecp2 2048 SPFEM_TILE_SORT=1
run pf150 vecp2 2048 SPFEM_TILE_PREFETCH=150
run pf600 vecp2 2048 SPFEM_TILE_PREFETCH=600
run minb4 vecp2 2048 SPFEM_TILE_MINB=4 SPFEM_TILE_E=48
run e48_t192 vecp2 2048 SPFEM_TILE_E=48 SPFEM_TILE_THREADS=192
run stokesb stokesb 2048 X=1
run stokesb_e64 stokesb 2048 SPFEM_TILE_E=64
run stokesb_e128 stokesb 2048 SPFEM_TILE_E=128
python - <<PY
import json
for l in open('gpurun_out/r02_c32_sweep.jsonl'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'),'thr',t.get('threads'), d.get('error','')[:80])
PY
