#!/bin/bash
# kernel variant for meshes of parallelograms only (no loop branch, 80 / 45 registers); the two new
# full-size property tests (config 4 TetP2 203^3, config 5 divergence block)
mkdir -p gpurun_out
OUT=gpurun_out/r02_c49_sweep.jsonl
run () { name=$1; kind=$2; size=$3; shift 3; echo "## $name $*" >> $OUT; env "$@" timeout 900 python tools/scale_probe.py $kind $size >> $OUT 2>> gpurun_out/r02_c49_sweep.err; }
rm -f $OUT
timeout 900 python -m pytest tests/test_quad_loop.py -m gpu -x -q > gpurun_out/r02_c49_quad_tests.log 2>&1; tail -n 2 gpurun_out/r02_c49_quad_tests.log
run q2 q2 11 X=1
run q2_mixed_kernel q2 11 SPFEM_QPARA=0
run q2_minb2 q2 11 SPFEM_TILE_MINB_AFFINE=2
run q2_np1 q2 11 SPFEM_TILE_NPART=1
run q2_np1_e96 q2 11 SPFEM_TILE_NPART=1 SPFEM_TILE_E=96
run q1 q1 11 X=1
run q1_mixed_kernel q1 11 SPFEM_QPARA=0
run q1_minb3 q1 11 SPFEM_TILE_MINB_AFFINE=3
python - <<PY
import json
for l in open('$OUT'):
    if l.startswith('##'): print(l.strip(), end=' -> ')
    else:
        d=json.loads(l); t=d.get('tile',{}); print(d.get('ms'), d.get('frac'), 'E',t.get('E'),'smem',t.get('smem'), d.get('error','')[:80])
PY
tail -n 3 gpurun_out/r02_c49_sweep.err
(time timeout 900 python -m pytest tests/test_parity_large.py -m gpu -x -q -k "config4 or config5") > gpurun_out/r02_c49_fullsize_tests.log 2>&1; tail -n 6 gpurun_out/r02_c49_fullsize_tests.log
