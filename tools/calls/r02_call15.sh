#!/bin/bash
# symmetric mirroring in the tile plan: parity + probes (mirror on / off)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c15_gpu_tests.log 2>&1
tail -n 4 gpurun_out/r02_c15_gpu_tests.log
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c15_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c15_sweep.jsonl 2>> gpurun_out/r02_c15_sweep.err; }
rm -f gpurun_out/r02_c15_sweep.jsonl
run vec vecp2 2048 X=1
run vec_nomirror vecp2 2048 SPFEM_TILE_MIRROR=0
run vec_sort1 vecp2 2048 SPFEM_TILE_SORT=1
run trip2 trip2 2048 X=1
run trip2_nomirror trip2 2048 SPFEM_TILE_MIRROR=0
run tet128 tetp2 128 X=1
run tet128_nomirror tetp2 128 SPFEM_TILE_MIRROR=0
run tet100 tetp2 100 X=1
run vec_full vecp2 3162 X=1
cut -c1-330 gpurun_out/r02_c15_sweep.jsonl
tail -n 3 gpurun_out/r02_c15_sweep.err
