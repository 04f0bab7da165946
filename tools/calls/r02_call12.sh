#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c12_sweep.jsonl; env "$@" timeout 300 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c12_sweep.jsonl 2>> gpurun_out/r02_c12_sweep.err; }
rm -f gpurun_out/r02_c12_sweep.jsonl
run vec_bucket vecp2 2048 SPFEM_TILE_SORT=bucket
run vec_bucket_half vecp2 2048 SPFEM_TILE_SORT=bucket SPFEM_TILE_X=X_HALF
run vec_sorted_half vecp2 2048 SPFEM_TILE_X=X_HALF
run trip2 trip2 2048 X=1
run trip2_half trip2 2048 SPFEM_TILE_X=X_HALF
run tet128 tetp2 128 X=1
run tet128_half tetp2 128 SPFEM_TILE_X=X_HALF
cut -c1-300 gpurun_out/r02_c12_sweep.jsonl
tail -n 3 gpurun_out/r02_c12_sweep.err
