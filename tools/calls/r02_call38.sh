#!/bin/bash
# compute-sanitizer memcheck over the kernels that changed this round (small meshes)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c38_plain.log 2>&1 || { tail -n 20 gpurun_out/r02_c38_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_quad_loop.py tests/test_field_fallback.py -m gpu -x -q \
    -k "golden or quadrature_loop_cuda or opaque" > gpurun_out/r02_c38_memcheck.log 2>&1
echo "memcheck exit code $?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_c38_memcheck.log
tail -n 12 gpurun_out/r02_c38_memcheck.log
