#!/usr/bin/env python
"""Regenerate ``spfem_b200/data/quadrature.json`` from the reference's own
tabulated rules (spfem/quadrature.py:64-114), evaluated through the patched
reference in ``oracle/_ref`` -- the numbers are never hand-copied.

Stored as ``repr``-exact float64 (JSON round-trips IEEE doubles exactly)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_ref
make_ref.build(quiet=True)
spfem = make_ref.import_ref()
from spfem.quadrature import get_quadrature_tri, get_quadrature_tet

out = {"tri": {}, "tet": {}}
for n in range(2, 20):
    X, W = get_quadrature_tri(n)
    out["tri"][str(n)] = {"X": X.tolist(), "W": W.tolist()}
for n in range(2, 5):
    X, W = get_quadrature_tet(n)
    out["tet"][str(n)] = {"X": X.tolist(), "W": W.tolist()}
dst = os.path.join(ROOT, "spfem_b200", "data", "quadrature.json")
with open(dst, "w") as fh:
    json.dump(out, fh)
print("wrote", dst, os.path.getsize(dst), "bytes")
