#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c24_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c24_gpu_tests.log
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c24_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c24_sweep.jsonl 2>> gpurun_out/r02_c24_sweep.err; }
rm -f gpurun_out/r02_c24_sweep.jsonl
run q2 q2 10 X=1
run q2_big q2 11 X=1
run q1 q1 11 X=1
cut -c1-330 gpurun_out/r02_c24_sweep.jsonl; tail -n 3 gpurun_out/r02_c24_sweep.err
