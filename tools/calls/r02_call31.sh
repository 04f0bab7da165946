#!/bin/bash
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c31_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c31_sweep.jsonl 2>> gpurun_out/r02_c31_sweep.err; }
rm -f gpurun_out/r02_c31_sweep.jsonl
run tet100 tetp2 100 X=1
run tet203 tetp2 203 X=1
run vec_full vecp2 3162 X=1
run stokesb stokesb 3162 X=1
run trip2 trip2 3162 X=1
run tetp1 tetp1 100 X=1
run q2 q2 11 X=1
cut -c1-330 gpurun_out/r02_c31_sweep.jsonl; tail -n 3 gpurun_out/r02_c31_sweep.err
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c31_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c31_gpu_tests.log
