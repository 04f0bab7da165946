#!/bin/bash
# tile kernel with interleaved warp lists: parity tests + probes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c10_gpu_tests.log 2>&1
tail -4 gpurun_out/r02_c10_gpu_tests.log
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c10_sweep.jsonl; env "$@" timeout 300 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c10_sweep.jsonl 2>> gpurun_out/r02_c10_sweep.err; }
rm -f gpurun_out/r02_c10_sweep.jsonl
run vec vecp2 2048 X=1
run vec_csr vecp2 2048 SPFEM_TILE_SORT=0
run vec_bucket vecp2 2048 SPFEM_TILE_SORT=bucket
run trip2 trip2 2048 X=1
run trip2_csr trip2 2048 SPFEM_TILE_SORT=0
run tet128 tetp2 128 X=1
run tet100 tetp2 100 X=1
run vec_full vecp2 3162 X=1
cut -c1-330 gpurun_out/r02_c10_sweep.jsonl
tail -3 gpurun_out/r02_c10_sweep.err
