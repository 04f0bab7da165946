#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's
# GPU minutes were spent and has only run on the host emulation / gloo so far.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -x
mkdir -p gpurun_out
# 1. the new GPU tests (GPU variants of the restated reference tests and chunked
#    assembly with the on-device merge: these passed on one B200 at the end of
#    round 1, profiles/r01_late_gpu_tests.log; the NCCL interface exchange at
#    world 2 has not run on GPUs yet)
python -m pytest tests/test_reference_module_tests.py tests/test_zz_nccl_exchange.py \
    tests/test_zz_chunked_gpu.py \
    -m gpu -q > gpurun_out/r02_new_gpu_tests.log 2>&1
tail -5 gpurun_out/r02_new_gpu_tests.log
# 2. halo vs exchange on the benchmark strip
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29611 tools/bench_exchange.py --refine 11 \
    > gpurun_out/r02_exchange_2gpu.json 2> gpurun_out/r02_exchange_2gpu.err
tail -2 gpurun_out/r02_exchange_2gpu.json
# 3. configs 3-5 over 2 GPUs, both modes
for mode in halo exchange; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29612 tools/bench_dist_configs.py --mode $mode \
    > gpurun_out/r02_dist_configs_${mode}_2gpu.json 2> gpurun_out/r02_dist_configs_${mode}.err
done
tail -3 gpurun_out/r02_dist_configs_halo_2gpu.json
# 4. the chunked path with the on-device merge at a size that is safe for the
#    host (result: 0.77e9 nonzeros = 12 GB of host arrays): MeshTri refine(11)
#    vector-P2 elasticity, 8.4 M triangles x 144 = 1.2e9 local entries, limit
#    lowered so that it runs in >= 3 chunks; compare with the single call.
#    (refine(12) = 4.8e9 entries needs ~50 GB of host memory for the result:
#    check `free -g` on the box before trying it.)
free -g | head -2
python - > gpurun_out/r02_chunked.log 2>&1 <<'PY'
import os, time, sys
sys.path.insert(0, "tests")
import numpy as np
import spfem_b200 as sp, cases
from spfem_b200.assembly import AssemblerElement
t0 = time.time(); m = sp.MeshTri(); m.refine(11); print("mesh", m.t.shape[1], time.time() - t0, flush=True)
a = AssemblerElement(m, sp.ElementH1Vec(sp.ElementTriP2()))
t0 = time.time(); A = a.iasm(cases.elasticity); dt = time.time() - t0
print("single call: nnz", A.nnz, "s", dt, flush=True)
a._engine = None
os.environ["SPFEM_MAX_ENTRIES"] = str(144 * m.t.shape[1] // 2)
t0 = time.time(); B = a.iasm(cases.elasticity); dt = time.time() - t0
print("chunked: nnz", B.nnz, "index dtype", B.indices.dtype, "s", dt,
      "elements/s", m.t.shape[1] / dt, flush=True)
assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
print("max |A - B| / max |A|", float(np.max(np.abs(A.data - B.data)) / np.max(np.abs(A.data))))
PY
tail -4 gpurun_out/r02_chunked.log
