#!/bin/bash
# Q1 / Q2 with the quadrature-point loop: parity + probes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_reference_known_answers.py tests/test_random_forms.py -m gpu -x -q -k "q1 or q2 or Q1 or Q2 or quad or Quad or random" > gpurun_out/r02_c20_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c20_gpu_tests.log
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c20_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c20_sweep.jsonl 2>> gpurun_out/r02_c20_sweep.err; }
rm -f gpurun_out/r02_c20_sweep.jsonl
run q2 q2 10 X=1
run q2_minb2 q2 10 SPFEM_TILE_MINB=2
run q2_minb2_np3 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_NPART=3
run q2_np4 q2 10 SPFEM_TILE_NPART=4
run q2_minb2_np1 q2 10 SPFEM_TILE_MINB=2 SPFEM_TILE_NPART=1
run q2_twokernel q2 10 SPFEM_FUSED=0
run q2_old q2 10 SPFEM_QLOOP=0
run q1 q1 11 X=1
run q1_minb2 q1 11 SPFEM_TILE_MINB=2
run q1_old q1 11 SPFEM_QLOOP=0
run q2_big q2 11 SPFEM_TILE_MINB=2
cut -c1-360 gpurun_out/r02_c20_sweep.jsonl
tail -n 3 gpurun_out/r02_c20_sweep.err
