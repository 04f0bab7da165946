#!/usr/bin/env python
"""Time variants of the vertex-fan kernel on ONE mesh / pattern (config 2 by
default): every variant is a set of environment switches read by the code
generator (kernel experiments ``SPFEM_FAN_X=flag,flag``, launch bounds, CTA size)
or by the plan builder (rows per tile).  The mesh, the assembler and the sorted
pattern are built once.

    python tools/fan_variants.py 11 [variant-file]

Each line of the variant file: ``name KEY=VALUE KEY=VALUE ...``.  Output: one
JSON line per variant (ms per assembly, fraction of the measured HBM peak).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import spfem_b200 as sp
from spfem_b200 import codegen
from spfem_b200.assembly import AssemblerElement
import cases

PEAK = 6558.1
DEFAULT = """
base
nocompute SPFEM_FAN_NOCOMPUTE=1
nostore SPFEM_FAN_X=X_NOSTORE
linear SPFEM_FAN_X=X_LINEAR
sameblob SPFEM_FAN_X=X_SAMEBLOB
warpstore SPFEM_FAN_X=WARPSTORE
"""
KEYS = ("SPFEM_FAN_NOCOMPUTE", "SPFEM_FAN_X", "SPFEM_FAN_T", "SPFEM_FAN_ROWS", "SPFEM_FAN_MINB8",
        "SPFEM_FAN_PREFETCH", "SPFEM_FAN_TINY", "SPFEM_FAN_ROWORDER", "SPFEM_RCP", "SPFEM_ADJUGATE",
        "SPFEM_FAN_NCOPY")


def sample(prep, steps=30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        prep.launch()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    text = open(sys.argv[2]).read() if len(sys.argv) > 2 else DEFAULT
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    m = sp.MeshTri()
    m.refine(level)
    nt, nv = m.t.shape[1], m.p.shape[1]
    a = AssemblerElement(m, sp.ElementTriP1())
    forms = {"lap": cases.lap2, "mass": cases.mass}
    ref, preps, results = {}, [], []
    for line in text.strip().splitlines():
        parts = line.split()
        if not parts or parts[0].startswith("#"):
            continue
        name, kv = parts[0], dict(x.split("=", 1) for x in parts[1:])
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(kv)
        codegen.FAN_THREADS = int(os.environ.get("SPFEM_FAN_T", "128"))
        eng = a._get_engine()
        for pat in eng.patterns.values():           # rows per tile / format may differ
            pat.__dict__.pop("_fan_plans", None)
        eng._tile_plans = [tp for tp in eng._tile_plans if tp.get("kind") != "vertex-fan"]
        res = {"variant": name, "env": kv}
        results.append(res)
        try:
            for fname, form in forms.items():
                prep = a.prepare_iasm(form)
                fp = prep.fp
                ns = fp["desc"][:, 4].long()
                fp["desc"][:, 6] = (torch.cumsum(ns, 0) - ns).to(torch.int32)
                for _ in range(3):
                    prep.launch()
                torch.cuda.synchronize()
                if not kv.get("SPFEM_FAN_X", "").startswith("X_") and "SPFEM_FAN_NOCOMPUTE" not in kv:
                    v = prep.values.clone()
                    if fname not in ref:
                        ref[fname] = v
                    res[fname + "_maxrel"] = float(((ref[fname] - v).abs().max() / ref[fname].abs().max()))
                res["format"], res["halo"], res["smem"] = fp["format"], round(fp["halo_factor"], 3), prep.smem
                preps.append((res, fname, prep, []))
        except Exception as exc:
            res["error"] = repr(exc)[:300]
    # interleaved rounds: clock / temperature drift hits every variant alike
    for _ in range(rounds):
        for res, fname, prep, ts in preps:
            ts.append(sample(prep))
    for res, fname, prep, ts in preps:
        nnz = prep.pattern.nnz
        balg = 12 * nt + 16 * nv + 8 * nnz
        ts.sort()
        res[fname + "_ms"] = round(ts[0], 4)
        res[fname + "_med"] = round(ts[len(ts) // 2], 4)
        res[fname + "_frac"] = round(balg / (ts[0] * 1e-3) / 1e9 / PEAK, 4)
    for res in results:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
