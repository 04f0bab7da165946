#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c23_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c23_sweep.jsonl 2>> gpurun_out/r02_c23_sweep.err; }
rm -f gpurun_out/r02_c23_sweep.jsonl
run q2_e96 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_E=96
run q2_e128 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_E=128
run q2_e96_np3 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_E=96 SPFEM_TILE_NPART=3
run q2_e128_np3 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_E=128 SPFEM_TILE_NPART=3
run q2_e64_np3 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_NPART=3
run q2_e96_np4 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_E=96 SPFEM_TILE_NPART=4
run q2_e96_t512 q2 10 SPFEM_TILE_MINB=1 SPFEM_TILE_E=96 SPFEM_TILE_NPART=4 SPFEM_TILE_T=512 SPFEM_TILE_THREADS=512
run q1_e512 q1 11 SPFEM_TABLE_SPACE=constant SPFEM_TILE_E=512
run q1_e128 q1 11 SPFEM_TABLE_SPACE=constant SPFEM_TILE_E=128
run q1_minb4 q1 11 SPFEM_TABLE_SPACE=constant SPFEM_TILE_MINB=4
cut -c1-330 gpurun_out/r02_c23_sweep.jsonl
tail -n 3 gpurun_out/r02_c23_sweep.err
