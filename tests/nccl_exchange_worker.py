# -*- coding: utf-8 -*-
"""Worker of tests/test_zz_nccl_exchange.py: the interface exchange of
``DistributedAssembler(mode="exchange")`` with the NCCL backend, one rank per
GPU, local assembly by the CUDA kernels.  Checked against the golden fixtures
and against the halo mode (which needs no collective)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

from spfem_b200.distributed import DistributedAssembler
import cases
import util


def main():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    for name in ("trip1_lap", "trip2_mass", "stokes_b", "tetp2_lap",
                 "trip1_load_sin_h"):
        case = cases.BY_NAME[name]
        m = util.mesh_for(case)
        eu = util.element(case.eu, case.ncu)
        ev = None
        if not (case.ev is None and case.ncv == case.ncu):
            ev = util.element(case.ev or case.eu, case.ncv)
        da = DistributedAssembler(m, eu, ev, mode="exchange")
        r1 = da.iasm(case.form)
        r2 = da.iasm(case.form)
        assert r1.values.is_cuda
        assert torch.equal(r1.values, r2.values), "exchange not reproducible"
        res = r1.gather(0)
        halo = DistributedAssembler(m, eu, ev, mode="halo").iasm(case.form).gather(0)
        if rank == 0:
            util.check(res, util.golden(case))
            if hasattr(res, "indptr"):
                cases.compare_csr(res, halo)
            else:
                cases.compare_vec(res, halo)
            print("nccl xchg ok", name, "world", world,
                  "interface entries", r1.xchg.interface_entries)
    # boundary-facet (Nitsche) terms in both modes
    for name in ("f_trip2_nitsche", "f_tetp2_bnd"):
        case = cases.BY_NAME[name]
        m = util.mesh_for(case)
        eu = util.element(case.eu, case.ncu)
        for mode in ("halo", "exchange"):
            res = DistributedAssembler(m, eu, None, mode=mode).fasm(case.form).gather(0)
            if rank == 0:
                util.check(res, util.golden(case))
                print("nccl facet ok", name, mode)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
