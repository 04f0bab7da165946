#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c21_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c21_sweep.jsonl 2>> gpurun_out/r02_c21_sweep.err; }
rm -f gpurun_out/r02_c21_sweep.jsonl
run q2_global_two q2 10 SPFEM_FUSED=0
run q2_const_two q2 10 SPFEM_FUSED=0 SPFEM_TABLE_SPACE=constant
run q2_global_tile q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_NPART=1
run q1_global q1 11 X=1
cut -c1-300 gpurun_out/r02_c21_sweep.jsonl
tail -n 3 gpurun_out/r02_c21_sweep.err
python tools/scale_probe.py q2 10 > gpurun_out/r02_c21_plain.log 2>&1 &&
SPFEM_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:spfem_elem -s 3 -c 1 -f -o gpurun_out/r02_c21_q2_elem env SPFEM_FUSED=0 python tools/scale_probe.py q2 10 > gpurun_out/r02_c21_ncu.log 2>&1
tail -n 2 gpurun_out/r02_c21_ncu.log
