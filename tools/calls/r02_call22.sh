#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c22_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c22_sweep.jsonl 2>> gpurun_out/r02_c22_sweep.err; }
rm -f gpurun_out/r02_c22_sweep.jsonl
run q2_two q2 10 SPFEM_FUSED=0
run q2_tile q2 10 X=1
run q2_tile_minb2 q2 10 SPFEM_TILE_MINB=2
run q2_tile_minb2_np1 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_NPART=1
run q2_tile_minb1_np1 q2 10 SPFEM_TILE_MINB=1 SPFEM_TILE_NPART=1
run q2_tile_np3 q2 10 SPFEM_TILE_NPART=3
run q2_tile_const q2 10 SPFEM_TILE_MINB=2 SPFEM_TABLE_SPACE=constant
run q1 q1 11 X=1
run q1_const q1 11 SPFEM_TABLE_SPACE=constant
cut -c1-300 gpurun_out/r02_c22_sweep.jsonl
tail -n 3 gpurun_out/r02_c22_sweep.err
