#!/usr/bin/env python
"""Device-resident throughput of the assembly kernels on the other element /
form combinations of BASELINE.json (TriP2, vector-P2 elasticity, Stokes B,
TetP1/P2) at sizes that fit one pattern call (< 2^31 local entries).

    python tools/bench_configs.py [--out gpurun_out/configs.json]

Prints one JSON object per configuration: elements/s for the fused tile kernel
and for the two-kernel route, algorithmic bytes and the
fraction of the measured HBM peak."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
import cases


def measure(prep, steps=20):
    for _ in range(3):
        prep.launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        prep.launch()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


CONFIGS = [
    ("TriP1 laplace", "MeshTri", 10, "TriP1", 1, None, 1, cases.lap2, 3),
    ("TriP1 load vector", "MeshTri", 10, "TriP1", 1, None, 1, cases.load1, 3),
    ("TriP2 laplace", "MeshTri", 9, "TriP2", 1, None, 1, cases.lap2, 6),
    ("vecP2 elasticity", "MeshTri", 9, "TriP2", 2, None, 2, cases.elasticity, 6),
    ("Stokes B (P1 x vecP2)", "MeshTri", 9, "TriP2", 2, "TriP1", 1, cases.stokes_b, 6),
    ("TetP1 laplace", "MeshTet", 6, "TetP1", 1, None, 1, cases.lap3, 4),
    ("TetP2 laplace", "MeshTet", 6, "TetP2", 1, None, 1, cases.lap3, 10),
    ("Q2 laplace", "MeshQuad", 9, "Q2", 1, None, 1, cases.lap2, 9),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="substring of the configuration name")
    ap.add_argument("--modes", default="fused,two_kernel,two_kernel_spatial")
    ap.add_argument("--force", action="store_true",
                    help="fused tile kernel even where the heuristic prefers two kernels")
    ap.add_argument("--refine-delta", type=int, default=0, dest="dref",
                    help="added to every configuration's refinement level")
    args = ap.parse_args()
    peak = 6558.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    out = []
    for name, mesh, ref, eu, ncu, ev, ncv, form, C in CONFIGS:
        if args.only and not any(name.startswith(o) for o in args.only.split(",")):
            continue
        m = getattr(sp, mesh)()
        m.refine(ref + args.dref)
        e_u = getattr(sp.element, "Element" + eu)()
        e_u = sp.ElementH1Vec(e_u) if ncu > 1 else e_u
        if ev is None and ncv == ncu:
            a = AssemblerElement(m, e_u)
        else:
            e_v = getattr(sp.element, "Element" + (ev or eu))()
            e_v = sp.ElementH1Vec(e_v) if ncv > 1 else e_v
            a = AssemblerElement(m, e_u, e_v)
        nt, nv, dim = m.t.shape[1], m.p.shape[1], m.p.shape[0]
        res = {"config": name, "nt": nt}
        for mode in args.modes.split(","):
            os.environ["SPFEM_FUSED"] = ("force" if args.force else "1") if mode == "fused" else "0"
            os.environ["SPFEM_LOCAL_LAYOUT"] = "spatial" if mode == "two_kernel_spatial" else "entry"
            a._engine = None
            t0 = time.perf_counter()
            prep = a.prepare_iasm(form)
            torch.cuda.synchronize()
            res[mode + "_setup_s"] = round(time.perf_counter() - t0, 2)
            if mode == "fused" and prep.n_launches != 1:
                res["fused_ms"] = None
                continue
            ms = measure(prep)
            nnz = prep.pattern.nnz if prep.plan.bilinear else prep.pattern.nrows
            balg = 4 * C * nt + 8 * dim * nv + 8 * nnz
            res["nnz"] = nnz
            res["B_alg_per_elem"] = round(balg / nt, 1)
            res[mode + "_ms"] = round(ms, 4)
            res[mode + "_elem_per_s"] = nt / (ms * 1e-3)
            res[mode + "_frac_of_hbm_peak"] = round(balg / (ms * 1e-3) / 1e9 / peak, 4)
            if mode == "fused":
                res["tile"] = {"E": prep.tp["E"], "ldl": prep.tp["ldl"], "threads": prep.threads,
                               "halo": round(prep.tp["halo_factor"], 3), "smem": prep.smem,
                               "direct": bool(prep.tp.get("direct", False)),
                               "mode": prep.tp.get("mode", "rows"),
                               "npart": getattr(prep.plan.spec, "npart", 1)}
            del prep
            torch.cuda.empty_cache()
        print(json.dumps(res))
        out.append(res)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
