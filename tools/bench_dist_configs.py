#!/usr/bin/env python
"""Weak-scaling timings of BASELINE configs 3-5 over N GPUs (one rank per GPU):
vector-P2 elasticity, TetP2 Laplace and the Stokes Taylor-Hood blocks A / B / C
plus the Nitsche boundary terms, through ``DistributedAssembler`` on the general
Morton partition of a replicated global mesh (halo mode: no data-path
collective; ``--mode exchange``: interface rows summed on their owners).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \\
        --master-addr 127.0.0.1 --master-port P tools/bench_dist_configs.py

The global mesh grows with N so that the work per GPU stays (about) fixed:
``--tri-refine`` / ``--tet-refine`` give the N = 1 mesh, one more uniform
refinement is a factor 4 (8 in 3-D), in between the unit square / cube is
repeated along x.  Per configuration one JSON object on rank 0: elements per
second over all ranks (CUDA events, max over ranks, 3 warm-up assemblies),
elements per rank with and without halo.  ``--host`` is a dry run of the
plumbing on CPUs (gloo, generated kernels compiled for the host; no timings
worth reading).  Not the contract bench (``bench.py``)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch
import torch.distributed as dist


def repeat_along_x(mesh, copies):
    """``copies`` translated copies of a unit-domain mesh glued along x
    (coincident vertices on the shared faces merged)."""
    import spfem_b200 as sp
    if copies == 1:
        return mesh
    p = np.hstack([mesh.p + np.eye(mesh.p.shape[0])[:, :1] * float(c) for c in range(copies)])
    nv = mesh.p.shape[1]
    t = np.hstack([mesh.t + c * nv for c in range(copies)])
    # merge duplicates: the coordinates of a refined unit mesh are exact dyadic
    # numbers, so equal points compare equal
    _, first, inv = np.unique(p.T, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    return type(mesh)(np.ascontiguousarray(p[:, first]), np.ascontiguousarray(inv[t]))


def sym_grad(du, dv):                       # Stokes A (learning/Example 1, cell 9)
    def eps(w, i, j):
        return 0.5 * (w[i][j] + w[j][i])
    return sum(eps(du, i, j) * eps(dv, i, j) for i in range(2) for j in range(2))


def vec_nitsche(u, v, du, dv, h, n):
    g = 10.0
    out = 0
    for i in range(2):
        out = out + g / h * u[i] * v[i]
        out = out - (du[i][0] * n[0] + du[i][1] * n[1]) * v[i]
        out = out - u[i] * (dv[i][0] * n[0] + dv[i][1] * n[1])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tri-refine", type=int, default=9, dest="tri")
    ap.add_argument("--tet-refine", type=int, default=5, dest="tet")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mode", default="halo", choices=["halo", "exchange"])
    ap.add_argument("--host", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if args.host:
        dist.init_process_group("gloo")
    else:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    import spfem_b200 as sp
    from spfem_b200.distributed import DistributedAssembler
    import cases
    hooks = {}
    if args.host:
        import hostemu
        hooks = dict(assemble=lambda a, f, **kw: hostemu.iasm(a, f, **kw),
                     assemble_facet=lambda a, f, **kw: hostemu.fasm(a, f, **kw))

    def sync():
        if not args.host:
            torch.cuda.synchronize()

    def timed(fn):
        for _ in range(1 if args.host else 3):
            fn()
        sync()
        dist.barrier()
        sync()
        steps = 1 if args.host else args.steps
        if args.host:
            t0 = time.perf_counter()
            fn()
            ms = 1e3 * (time.perf_counter() - t0)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], dtype=torch.float64, device="cpu" if args.host else "cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def grown(cls, refine, factor):
        """Unit mesh refined so that nt grows ~ world: whole refinements while
        they fit, then copies along x."""
        m = cls()
        extra = 0
        while factor ** (extra + 1) <= world:
            extra += 1
        m.refine(refine + extra)
        return repeat_along_x(m, world // factor ** extra)

    P2, P1 = sp.ElementTriP2, sp.ElementTriP1
    vecP2 = lambda: sp.ElementH1Vec(P2())
    tri = [("config 3: vector-P2 elasticity", vecP2, None, cases.elasticity, "iasm"),
           ("config 5: Stokes A (sym-grad, vector-P2)", vecP2, None, sym_grad, "iasm"),
           ("config 5: Stokes B (P1 x vector-P2)", vecP2, P1, cases.stokes_b, "iasm"),
           ("config 5: Stokes C (P1 mass)", P1, None, cases.mass, "iasm"),
           ("config 5: Nitsche boundary terms (vector-P2)", vecP2, None, vec_nitsche, "fasm")]
    tet = [("config 4: TetP2 Laplace", sp.ElementTetP2, None, cases.lap3, "iasm")]
    for cls, refine, factor, configs in ((sp.MeshTri, args.tri, 4, tri),
                                         (sp.MeshTet, args.tet, 8, tet)):
        configs = [c for c in configs if not args.only or args.only in c[0]]
        if not configs:
            continue
        t0 = time.perf_counter()
        mesh = grown(cls, refine, factor)
        t_mesh = time.perf_counter() - t0
        for name, eu, ev, form, kind in configs:
            t0 = time.perf_counter()
            da = DistributedAssembler(mesh, eu(), ev() if ev else None, mode=args.mode, **hooks)
            t_part = time.perf_counter() - t0
            if kind == "fasm":
                fn = lambda: da.fasm(form)
            elif args.host:
                fn = lambda: da.iasm(form)
            else:
                # staged once (trace, compile, pattern, plans); the timed call is
                # the kernel launch (+ the interface exchange in exchange mode)
                prep = da.prepare_iasm(form)
                if args.mode == "halo":
                    fn = prep.launch
                else:
                    def fn():
                        prep.launch()
                        return da._exchange(prep.result(to_host=False), prep.plan.bilinear)
            ms = timed(fn)
            nt = mesh.t.shape[1]
            loc = torch.tensor([da.part.mesh.t.shape[1], da.part.n_own_elements],
                               dtype=torch.int64, device="cpu" if args.host else "cuda")
            dist.all_reduce(loc, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({
                    "config": name, "n_gpus": world, "mode": args.mode, "kind": kind,
                    "elements": int(nt), "ms_per_assembly": ms,
                    "elements_per_s": nt / (ms * 1e-3) if kind == "iasm" else None,
                    "max_local_elements": int(loc[0]), "max_own_elements": int(loc[1]),
                    "mesh_s": round(t_mesh, 2), "partition_s": round(t_part, 2),
                    "host_dry_run": bool(args.host)}), flush=True)
            del da
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
