#!/usr/bin/env python
"""FP64 peak of this B200, measured: independent DFMA chains
in registers, no memory traffic.  Compiled by NVRTC for sm_100a through the
C-ABI like the generated kernels.  Writes ``MEASURED_FP64.json`` next to the
driver's ``MEASURED_PEAKS.json`` (which it never touches).

    python tools/fp64_peak.py [out.json]
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spfem_b200 import _lib, codegen
from spfem_b200.engine import get_kernel

SRC = codegen.PRELUDE + r"""
#define CHAINS 8
extern "C" __global__ void __launch_bounds__(256) spfem_fp64_peak(const SpfemArgs P)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double x = P.params[0], y = P.params[1];
    double a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) a[k] = (double)(tid + k) * 1e-9;
    for (long long i = 0; i < P.ldo; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) a[k] = fma(a[k], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += a[k];
    if (tid < P.n) P.out[tid] = s;
}
"""


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "MEASURED_FP64.json")
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    props = torch.cuda.get_device_properties(dev)
    kernel = get_kernel(SRC, "spfem_fp64_peak")
    nthreads = props.multi_processor_count * 2048
    iters = 8192
    out = torch.empty(nthreads, dtype=torch.float64, device=dev)
    a = _lib.SpfemArgs()
    a.n, a.ldo = nthreads, iters
    a.params[0], a.params[1] = 0.999999, 1e-7
    a.out = out.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(3):
        _lib.check(lib.spfem_launch(kernel, C.byref(a), 256, st))
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spfem_launch(kernel, C.byref(a), 256, st))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * 8 * iters * nthreads
    res = {
        "fp64_tflops": round(flops / (best * 1e-3) / 1e12, 2),
        "gpu_name": props.name, "sms": props.multi_processor_count,
        "how": "8 independent DFMA chains per thread, %d iterations, %d threads (2048 per SM), "
               "best of 10 launches, CUDA events; NVRTC sm_100a through libspfem_b200.so"
               % (iters, nthreads),
        "ms": round(best, 4),
        "dfma_per_clk_per_sm_at_1965MHz": round(flops / 2 / (best * 1e-3) / props.multi_processor_count / 1.965e9, 1),
        "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
    }
    print(json.dumps(res))
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
