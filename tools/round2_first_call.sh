#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's
# GPU minutes were spent and has only run on the host emulation / gloo so far.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -x
mkdir -p gpurun_out
# 1. the new GPU tests (GPU variants of the restated reference tests, chunked
#    assembly with the on-device merge, NCCL interface exchange at world 2)
python -m pytest tests/test_reference_module_tests.py tests/test_zz_nccl_exchange.py \
    "tests/test_gpu_parity.py::test_chunked_assembly_beyond_one_pattern_call" \
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
# 4. config 3 at (nearly) full size on one GPU through the chunked path:
#    MeshTri refine(11) vector-P2 = 8.4 M triangles x 144 entries = 1.2e9 (one call);
#    refine(12) = 33.5 M x 144 = 4.8e9 entries -> chunks, on-device merge
python - > gpurun_out/r02_chunked_full.log 2>&1 <<'PY'
import time, sys
sys.path.insert(0, "tests")
import spfem_b200 as sp, cases
from spfem_b200.assembly import AssemblerElement
t0 = time.time(); m = sp.MeshTri(); m.refine(12); print("mesh", m.t.shape[1], time.time() - t0, flush=True)
a = AssemblerElement(m, sp.ElementH1Vec(sp.ElementTriP2()))
t0 = time.time(); A = a.iasm(cases.elasticity); dt = time.time() - t0
print("chunked vecP2 elasticity: nnz", A.nnz, "index dtype", A.indices.dtype, "s", dt,
      "elements/s", m.t.shape[1] / dt, flush=True)
PY
tail -3 gpurun_out/r02_chunked_full.log
