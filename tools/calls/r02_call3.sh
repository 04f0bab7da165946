#!/bin/bash
# Round 2, call 3 (1 GPU): parity of the tile-kernel variants, tuning sweep, ncu of the block tile kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r02_c3_parity.log 2>&1
tail -5 gpurun_out/r02_c3_parity.log
run() {  # name kind size env...
  name=$1; kind=$2; size=$3; shift 3
  echo "## $name $*" >> gpurun_out/r02_c3_sweep.jsonl
  env "$@" timeout 300 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c3_sweep.jsonl 2>> gpurun_out/r02_c3_sweep.err
}
run default vecp2 1024 X=1
run scalar_slots vecp2 1024 SPFEM_TILE_BLOCKS=0
run E48 vecp2 1024 SPFEM_TILE_E=48
run E96 vecp2 1024 SPFEM_TILE_E=96
run E128 vecp2 1024 SPFEM_TILE_E=128
run minb3 vecp2 1024 SPFEM_TILE_T=256 SPFEM_TILE_MINB=3
run minb4 vecp2 1024 SPFEM_TILE_T=256 SPFEM_TILE_MINB=4
run t128 vecp2 1024 SPFEM_TILE_THREADS=128
run t128minb6 vecp2 1024 SPFEM_TILE_THREADS=128 SPFEM_TILE_T=128 SPFEM_TILE_MINB=6
run t512 vecp2 1024 SPFEM_TILE_THREADS=512
run bucket vecp2 1024 SPFEM_TILE_SORT=bucket
run sort1 vecp2 1024 SPFEM_TILE_SORT=1
run pf0 vecp2 1024 SPFEM_TILE_PREFETCH=0
run pf600 vecp2 1024 SPFEM_TILE_PREFETCH=600
run full vecp2 3162 X=1
run full_minb3 vecp2 3162 SPFEM_TILE_T=256 SPFEM_TILE_MINB=3
run tet_default tetp2 64 X=1
run tet_E32 tetp2 64 SPFEM_TILE_E=32
run tet_E48 tetp2 64 SPFEM_TILE_E=48
run tet_E96 tetp2 64 SPFEM_TILE_E=96
run tet_dense tetp2 64 SPFEM_TILE_MODE=dense SPFEM_TILE_E=32
run tet_minb2 tetp2 64 SPFEM_TILE_T=256 SPFEM_TILE_MINB=2
run tet_t512 tetp2 64 SPFEM_TILE_THREADS=512
run tet100 tetp2 100 X=1
run tet160 tetp2 160 X=1
run tet203 tetp2 203 X=1
run trip2 trip2 1024 X=1
run trip2_E64 trip2 1024 SPFEM_TILE_E=64
run trip2_E128 trip2 1024 SPFEM_TILE_E=128
run trip2_minb3 trip2 1024 SPFEM_TILE_T=256 SPFEM_TILE_MINB=3
cat gpurun_out/r02_c3_sweep.jsonl | cut -c1-420
# ncu: block tile kernel, vec-P2 n=1024
ncu --set full --clock-control none --import-source on -k regex:spfem_elem_tile -s 3 -c 1 \
    -o gpurun_out/r02_vecp2_tile_blocks python tools/scale_probe.py vecp2 1024 > gpurun_out/r02_c3_ncu.log 2>&1
tail -3 gpurun_out/r02_c3_ncu.log
ncu --set full --clock-control none --import-source on -k regex:spfem_elem_tile -s 3 -c 1 \
    -o gpurun_out/r02_tetp2_tile python tools/scale_probe.py tetp2 64 > gpurun_out/r02_c3_ncu2.log 2>&1
tail -3 gpurun_out/r02_c3_ncu2.log
