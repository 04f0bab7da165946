#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c30_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c30_sweep.jsonl 2>> gpurun_out/r02_c30_sweep.err; }
rm -f gpurun_out/r02_c30_sweep.jsonl
run np1_e64 tetp2 100 SPFEM_TILE_NPART=1
run np1_e72 tetp2 100 SPFEM_TILE_NPART=1 SPFEM_TILE_E=72
run np1_e80 tetp2 100 SPFEM_TILE_NPART=1 SPFEM_TILE_E=80
run np1_e88 tetp2 100 SPFEM_TILE_NPART=1 SPFEM_TILE_E=88
run np1_e48_minb3 tetp2 100 SPFEM_TILE_NPART=1 SPFEM_TILE_E=48 SPFEM_TILE_MINB=3
run np1_e56_minb3 tetp2 100 SPFEM_TILE_NPART=1 SPFEM_TILE_E=56 SPFEM_TILE_MINB=3
run np1_e64_t192 tetp2 100 SPFEM_TILE_NPART=1 SPFEM_TILE_THREADS=192
run np1_e128 tetp2 128 SPFEM_TILE_NPART=1
run np1_203 tetp2 203 SPFEM_TILE_NPART=1
run vec_np1 vecp2 2048 SPFEM_TILE_NPART=1
run trip2_np2 trip2 2048 SPFEM_TILE_NPART=2
cut -c1-330 gpurun_out/r02_c30_sweep.jsonl; tail -n 3 gpurun_out/r02_c30_sweep.err
