#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c25_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c25_sweep.jsonl 2>> gpurun_out/r02_c25_sweep.err; }
rm -f gpurun_out/r02_c25_sweep.jsonl
run vec vecp2 2048 X=1
run vec2 vecp2 2048 X=1
run stokesb stokesb 2048 X=1
run vec_full vecp2 3162 X=1
run q2 q2 11 X=1
run q1 q1 11 X=1
cut -c1-300 gpurun_out/r02_c25_sweep.jsonl; tail -n 3 gpurun_out/r02_c25_sweep.err
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_large.py tests/test_quad_loop.py -m gpu -x -q > gpurun_out/r02_c25_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c25_gpu_tests.log
