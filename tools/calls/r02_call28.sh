#!/bin/bash
# majority row ownership in the tile plan: probes (owner rule 1 = default vs 0 = lowest tile)
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c28_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c28_sweep.jsonl 2>> gpurun_out/r02_c28_sweep.err; }
rm -f gpurun_out/r02_c28_sweep.jsonl
run tet100 tetp2 100 X=1
run tet100_low tetp2 100 SPFEM_TILE_OWNER=0
run tet128 tetp2 128 X=1
run tet128_low tetp2 128 SPFEM_TILE_OWNER=0
run vec vecp2 2048 X=1
run vec_low vecp2 2048 SPFEM_TILE_OWNER=0
run trip2 trip2 2048 X=1
run trip2_low trip2 2048 SPFEM_TILE_OWNER=0
run tet203 tetp2 203 X=1
run vec_full vecp2 3162 X=1
run tet100_e80 tetp2 100 SPFEM_TILE_E=80
run tet100_e96 tetp2 100 SPFEM_TILE_E=96
cut -c1-420 gpurun_out/r02_c28_sweep.jsonl; tail -n 3 gpurun_out/r02_c28_sweep.err
