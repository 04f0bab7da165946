#!/bin/bash
# loop-free parallelogram branch of the quadrilateral kernels: Q2 / Q1 with and without it, quad parity tests
mkdir -p gpurun_out
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> gpurun_out/r02_c44_sweep.jsonl; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> gpurun_out/r02_c44_sweep.jsonl 2>> gpurun_out/r02_c44_sweep.err; }
rm -f gpurun_out/r02_c44_sweep.jsonl
run q2 q2 11 X=1
run q2_general q2 11 SPFEM_QAFFINE=0
run q1 q1 11 X=1
run q1_general q1 11 SPFEM_QAFFINE=0
python - <<PY
import json
for l in open('gpurun_out/r02_c44_sweep.jsonl'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'), d.get('error','')[:80])
PY
tail -n 3 gpurun_out/r02_c44_sweep.err
timeout 900 python -m pytest tests/test_quad_loop.py tests/test_gpu_parity.py tests/test_reference_known_answers.py -m gpu -x -q -k "quad or q1 or q2 or Q1 or Q2 or parallelogram" > gpurun_out/r02_c44_gpu_tests.log 2>&1; tail -n 2 gpurun_out/r02_c44_gpu_tests.log
