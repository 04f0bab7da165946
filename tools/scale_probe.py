#!/usr/bin/env python
"""Where set-up time and memory go as the mesh grows (one configuration family
per call): mesh build, Dofnum, pattern, plan, kernel time, roofline fraction.

    python tools/scale_probe.py vecp2 512 1024 2048
    python tools/scale_probe.py tetp2 32 64 100
    python tools/scale_probe.py trip1 11 12 13        (refine levels)
"""
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
from spfem_b200 import meshgen
from spfem_b200.assembly import AssemblerElement
import cases

PEAK = 6558.1


def measure(prep, steps=10):
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


def run(kind, size):
    res = {"kind": kind, "size": size}
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    if kind == "vecp2":
        p, t = meshgen.grid_tri(size)
        m = sp.MeshTri(p, t, validate=False)
        elem, form, C = sp.ElementH1Vec(sp.ElementTriP2()), cases.elasticity, 6
    elif kind == "stokesb":
        p, t = meshgen.grid_tri(size)
        m = sp.MeshTri(p, t, validate=False)
        elem, form, C = sp.ElementH1Vec(sp.ElementTriP2()), cases.stokes_b, 6
        elem_v = sp.ElementTriP1()
    elif kind == "trip2":
        p, t = meshgen.grid_tri(size)
        m = sp.MeshTri(p, t, validate=False)
        elem, form, C = sp.ElementTriP2(), cases.lap2, 6
    elif kind == "tetp1":
        p, t = meshgen.grid_tet(size)
        m = sp.MeshTet(p, t, validate=False)
        elem, form, C = sp.ElementTetP1(), cases.lap3, 4
    elif kind == "tetp2":
        p, t = meshgen.grid_tet(size)
        m = sp.MeshTet(p, t, validate=False)
        elem, form, C = sp.ElementTetP2(), cases.lap3, 10
    elif kind in ("q1", "q2"):
        m = sp.MeshQuad()
        m.refine(size)
        elem, form, C = (sp.ElementQ1() if kind == "q1" else sp.ElementQ2()), cases.lap2, 4
    elif kind == "trip1":
        m = sp.MeshTri()
        m.refine(size)
        elem, form, C = sp.ElementTriP1(), cases.lap2, 3
    else:
        raise SystemExit("unknown kind")
    res["mesh_s"] = round(time.perf_counter() - t0, 2)
    nt, nv, dim = m.t.shape[1], m.p.shape[1], m.p.shape[0]
    res["nt"] = nt
    t0 = time.perf_counter()
    a = AssemblerElement(m, elem, locals().get('elem_v'))
    res["dofnum_s"] = round(time.perf_counter() - t0, 2)
    t0 = time.perf_counter()
    prep = a.prepare_iasm(form)
    torch.cuda.synchronize()
    res["prepare_s"] = round(time.perf_counter() - t0, 2)
    res["stats"] = {k: round(v, 3) for k, v in a._get_engine().stats.items()}
    res["route"] = type(prep).__name__
    ms = measure(prep)
    nnz = prep.pattern.nnz
    balg = 4 * C * nt + 8 * dim * nv + 8 * nnz
    res.update(nnz=nnz, ms=round(ms, 4), elem_per_s=nt / (ms * 1e-3),
               frac=round(balg / (ms * 1e-3) / 1e9 / PEAK, 4),
               peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))
    if hasattr(prep, "tp"):
        res["tile"] = {k: prep.tp.get(k) for k in ("E", "halo_factor", "ntiles", "compact", "sort_slots",
                                                   "blob_bytes", "bsv", "bsu", "ldl", "max_ne",
                                                   "scratch_peak_bytes")}
        res["tile"]["batches"] = len(prep.tp.get("batches", [])) or None
        res["tile"]["smem"] = prep.smem
        res["tile"]["threads"] = prep.threads
    return res


def main():
    kind = sys.argv[1]
    for s in sys.argv[2:]:
        try:
            print(json.dumps(run(kind, int(s))), flush=True)
        except Exception as exc:                     # keep going: report the failure
            print(json.dumps({"kind": kind, "size": int(s), "error": repr(exc)[:300]}), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
