#!/usr/bin/env python
"""Where the time of one end-to-end ``a.iasm(form)`` (host CSR out) goes on the
benchmark mesh: upload, plan lookup, kernel, D2H of the values, host copies of
the cached pattern."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
from spfem_b200 import engine as E

m = sp.MeshTri(); m.refine(int(sys.argv[1]) if len(sys.argv) > 1 else 11)
a = AssemblerElement(m, sp.ElementTriP1())
f = lambda du, dv: du[0] * dv[0] + du[1] * dv[1]
g = lambda u, v: u * v
for _ in range(4):
    A = a.iasm(f); B = a.iasm(g)
eng = a._get_engine()
def T(fn, n=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
print("upload_mesh      %.2f ms" % T(eng.upload_mesh)[0])
print("plan_iasm        %.2f ms" % T(lambda: a.plan_iasm(f))[0])
plan = a.plan_iasm(f)
print("prepare_interior %.2f ms" % T(lambda: eng.prepare_interior(plan, None, None))[0])
prep = eng.prepare_interior(plan, None, None)
print("launch+sync      %.2f ms" % T(prep.launch)[0])
vals = prep.launch()
print("D2H values       %.2f ms (%.0f MB)" % (T(lambda: E._to_host(vals))[0], vals.numel() * 8 / 1e6))
ip, ix = prep.pattern.host()
print("copy indices     %.2f ms (%.0f MB)" % (T(lambda: E._host_copy(ix))[0], ix.nbytes / 1e6))
print("copy indptr      %.2f ms" % T(lambda: E._host_copy(ip))[0])
d = E.DeviceCSR(prep.pattern, vals, (a.dofnum_v.N, a.dofnum_u.N))
print("to_scipy         %.2f ms" % T(d.to_scipy)[0])
print("iasm             %.2f ms" % T(lambda: a.iasm(f))[0])
def two():
    A = a.iasm(f); B = a.iasm(g); return A, B
print("2 x iasm (held)  %.2f ms" % T(two)[0])
