#!/bin/bash
# Round 2, call 4 (1 GPU): full GPU suite, fan-kernel sweep (tiles per CTA), bench with the configs array
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c4_gpu_tests.log 2>&1
tail -8 gpurun_out/r02_c4_gpu_tests.log
run() {  # name kind size env...
  name=$1; kind=$2; size=$3; shift 3
  echo "## $name $*" >> gpurun_out/r02_c4_sweep.jsonl
  env "$@" timeout 300 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c4_sweep.jsonl 2>> gpurun_out/r02_c4_sweep.err
}
for K in 1 2 3 4 6 8 16 32; do
  run fanK$K trip1 11 SPFEM_FAN_K=$K
  run fanK${K}_minb9 trip1 11 SPFEM_FAN_K=$K SPFEM_FAN_MINB8=9
done
run fanK4_minb8 trip1 11 SPFEM_FAN_K=4 SPFEM_FAN_MINB8=8
run fanK4_minb10 trip1 11 SPFEM_FAN_K=4 SPFEM_FAN_MINB8=10
run fanK4_rows64 trip1 11 SPFEM_FAN_K=8 SPFEM_FAN_ROWS=64 SPFEM_FAN_T=64 SPFEM_FAN_MINB8=16
run fanK4_rows256 trip1 11 SPFEM_FAN_K=4 SPFEM_FAN_ROWS=256 SPFEM_FAN_T=256 SPFEM_FAN_MINB8=4
run fanK4_pf0 trip1 11 SPFEM_FAN_K=4 SPFEM_FAN_PREFETCH=0
run vec_full vecp2 3162 X=1
run vec_1024 vecp2 1024 X=1
run vec_1024_nosort vecp2 1024 SPFEM_TILE_SORT=0
run stokesb_like trip2 3162 X=1
run tet_100 tetp2 100 X=1
cut -c1-330 gpurun_out/r02_c4_sweep.jsonl
(time python bench.py --steps 100) > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err
tail -c 6000 gpurun_out/r02_c4_bench.json; tail -5 gpurun_out/r02_c4_bench.err
