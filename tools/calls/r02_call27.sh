#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_large.py -m gpu -x -q -k "fan or trip1 or config2 or reproducible or half_a_million" > gpurun_out/r02_c27_gpu_tests.log 2>&1; tail -n 3 gpurun_out/r02_c27_gpu_tests.log
for i in 1 2; do python bench.py --configs none --no-cpu-baseline > gpurun_out/r02_c27_bench$i.json 2> gpurun_out/r02_c27_bench$i.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c27_bench$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernels'])
PY
done
SPFEM_FAN_ROWS=128 python bench.py --configs none --no-cpu-baseline > gpurun_out/r02_c27_bench_r128.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c27_bench_r128.json').read().strip().splitlines()[-1])
print('rows128', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernels'])
PY
