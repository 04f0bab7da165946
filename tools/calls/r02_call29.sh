#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c29_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c29_sweep.jsonl 2>> gpurun_out/r02_c29_sweep.err; }
rm -f gpurun_out/r02_c29_sweep.jsonl
run tet100 tetp2 100 X=1
run tet100_e48 tetp2 100 SPFEM_TILE_E=48
run tet100_e48_minb3 tetp2 100 SPFEM_TILE_E=48 SPFEM_TILE_MINB=3
run tet100_e56 tetp2 100 SPFEM_TILE_E=56
run tet100_e72 tetp2 100 SPFEM_TILE_E=72
run tet100_np3 tetp2 100 SPFEM_TILE_NPART=3
run tet100_np4 tetp2 100 SPFEM_TILE_NPART=4
run tet100_np1 tetp2 100 SPFEM_TILE_NPART=1
run tet100_t512 tetp2 100 SPFEM_TILE_THREADS=512 SPFEM_TILE_T=512 SPFEM_TILE_MINB=1
run tet100_dense tetp2 100 SPFEM_TILE_MODE=dense SPFEM_TILE_E=32
run tetp1 tetp1 100 X=1
cut -c1-330 gpurun_out/r02_c29_sweep.jsonl; tail -n 3 gpurun_out/r02_c29_sweep.err
