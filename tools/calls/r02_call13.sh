#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02_c13_gpu_tests.log 2>&1
tail -25 gpurun_out/r02_c13_gpu_tests.log
python tools/fp64_peak.py > gpurun_out/r02_c13_fp64.json 2> gpurun_out/r02_c13_fp64.err; cat gpurun_out/r02_c13_fp64.json; tail -n 2 gpurun_out/r02_c13_fp64.err
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c13_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c13_sweep.jsonl 2>> gpurun_out/r02_c13_sweep.err; }
rm -f gpurun_out/r02_c13_sweep.jsonl
run q2 q2 10 X=1
run q2_force q2 10 SPFEM_FUSED=force
run q1 q1 11 X=1
cut -c1-400 gpurun_out/r02_c13_sweep.jsonl; tail -n 3 gpurun_out/r02_c13_sweep.err
(time python bench.py) > gpurun_out/r02_c13_bench.json 2> gpurun_out/r02_c13_bench.err
tail -n 5 gpurun_out/r02_c13_bench.err; cut -c1-1500 gpurun_out/r02_c13_bench.json
