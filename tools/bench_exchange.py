# -*- coding: utf-8 -*-
"""Halo mode vs interface-exchange mode of the multi-GPU assembly on the
weak-scaling strip mesh (one ``MeshTri().refine(R)`` unit square per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \\
        --master-addr 127.0.0.1 --master-port P tools/bench_exchange.py --refine 11

Per mode: ms per assembly (CUDA events on the launching stream, max over ranks,
after 3 warm-up assemblies) of the Laplace form, the number of interface
entries a rank sends, and a parity check of the two gathered matrices at a
small refinement.  One JSON line on rank 0.  Not the contract bench (bench.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch
import torch.distributed as dist


def lap2(du, dv):
    return du[0] * dv[0] + du[1] * dv[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=11)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    import spfem_b200 as sp
    from spfem_b200.distributed import DistributedAssembler, strip_part
    import cases

    # parity of the two modes (small mesh, gathered on rank 0)
    got = {}
    for mode in ("halo", "exchange"):
        part = strip_part(rank, world, 4, halo=(mode == "halo"))
        da = DistributedAssembler(None, sp.ElementTriP1(), part=part, mode=mode)
        got[mode] = da.iasm(lap2).gather(0)
    if rank == 0:
        cases.compare_csr(got["exchange"], got["halo"])

    out = {"refine": args.refine, "n_gpus": world, "steps": args.steps}
    for mode in ("halo", "exchange"):
        part = strip_part(rank, world, args.refine, halo=(mode == "halo"))
        da = DistributedAssembler(None, sp.ElementTriP1(), part=part, mode=mode)
        for _ in range(3):
            res = da.iasm(lap2)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            res = da.iasm(lap2)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda",
                          dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[mode] = {"ms_per_assembly": float(ms.item()),
                     "local_elements": int(part.mesh.t.shape[1])}
        if mode == "exchange":
            out[mode]["interface_entries_sent"] = res.xchg.interface_entries
            out[mode]["owned_nnz"] = res.xchg.nnz
        del da, res
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
