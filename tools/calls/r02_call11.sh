#!/bin/bash
# ncu of the tile kernel with interleaved warp lists (vector-P2 2-D, TetP2 3-D)
mkdir -p gpurun_out
export SPFEM_TILE_SORT=bucket
python tools/scale_probe.py vecp2 1024 > gpurun_out/r02_c11_plain_vec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spfem_elem_tile -s 3 -c 1 -f -o gpurun_out/r02_c11_vecp2_tile python tools/scale_probe.py vecp2 1024 > gpurun_out/r02_c11_ncu_vec.log 2>&1
unset SPFEM_TILE_SORT
python tools/scale_probe.py tetp2 64 > gpurun_out/r02_c11_plain_tet.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spfem_elem_tile -s 3 -c 1 -f -o gpurun_out/r02_c11_tetp2_tile python tools/scale_probe.py tetp2 64 > gpurun_out/r02_c11_ncu_tet.log 2>&1
cat gpurun_out/r02_c11_plain_vec.log gpurun_out/r02_c11_plain_tet.log | cut -c1-300
tail -3 gpurun_out/r02_c11_ncu_vec.log gpurun_out/r02_c11_ncu_tet.log
