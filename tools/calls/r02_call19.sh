#!/bin/bash
# bank-friendly numbering of the distinct local values + row-major UT index: probes
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c19_sweep.jsonl; env "$@" timeout 600 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c19_sweep.jsonl 2>> gpurun_out/r02_c19_sweep.err; }
rm -f gpurun_out/r02_c19_sweep.jsonl
run vec vecp2 2048 X=1
run vec_colorder vecp2 2048 SPFEM_UNIQ_ORDER=col
run trip2 trip2 2048 X=1
run trip2_colorder trip2 2048 SPFEM_UNIQ_ORDER=col
run tet128 tetp2 128 X=1
run tet128_colorder tetp2 128 SPFEM_UNIQ_ORDER=col
run fan trip1 11 X=1
cut -c1-300 gpurun_out/r02_c19_sweep.jsonl
tail -n 3 gpurun_out/r02_c19_sweep.err
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_parity_large.py -m gpu -x -q > gpurun_out/r02_c19_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c19_gpu_tests.log
