#!/usr/bin/env python
"""Generate the golden fixtures ``tests/golden/<case>.npz`` by running the
REFERENCE ITSELF (``/root/reference/spfem``, mechanically Python-3 patched by
``oracle/make_ref.py`` into ``oracle/_ref``) on the cases of ``tests/cases.py``.

Run in the build container only (needs ``/root/reference``):
    python tests/golden/make_golden.py
Each fixture stores the reference's CSR ``indptr / indices / data`` (or the
load vector) plus the mesh size; the GPU box never needs the reference."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import make_ref  # noqa: E402

make_ref.build(quiet=True)
spfem = make_ref.import_ref()
import spfem.mesh as rmesh  # noqa: E402
import spfem.element as relem  # noqa: E402
import spfem.assembly as rasm  # noqa: E402
import cases  # noqa: E402


def ref_element(name, ncomp):
    if name == "TriDG1":
        e = relem.ElementTriDG(relem.ElementTriP1())
    elif name.startswith("TriPp"):
        e = relem.ElementTriPp(int(name[5:]))
    else:
        e = getattr(relem, "Element" + name)()
    return relem.ElementH1Vec(e) if ncomp > 1 else e


def main():
    for c in cases.CASES:
        m = getattr(rmesh, c.mesh)()
        m.refine(c.refine)
        eu = ref_element(c.eu, c.ncu)
        if c.ev is None and c.ncv == c.ncu:
            a = rasm.AssemblerElement(m, eu)
        else:
            a = rasm.AssemblerElement(m, eu, ref_element(c.ev or c.eu, c.ncv))
        kw = cases.case_inputs(c, m, a.dofnum_u.N)
        out = a.iasm(c.form, **kw) if c.kind == "iasm" else a.fasm(c.form, **kw)
        path = os.path.join(HERE, c.name + ".npz")
        if hasattr(out, "indptr"):
            np.savez_compressed(path, kind="csr", shape=np.array(out.shape),
                                indptr=out.indptr, indices=out.indices,
                                data=out.data, nt=m.t.shape[1])
            print("%-22s csr %s nnz %d %s" % (c.name, out.shape, out.nnz, out.indices.dtype))
        else:
            np.savez_compressed(path, kind="vec", data=out, nt=m.t.shape[1])
            print("%-22s vec %d" % (c.name, out.shape[0]))
    # known answer of the reference's own test (spfem/test_asm.py:145-163)
    m = rmesh.MeshTri()
    m.refine(5)
    a = rasm.AssemblerElement(m, relem.ElementTriP1())
    A = a.iasm(cases.lap2)
    b = a.iasm(cases.load1)
    import spfem.utils as rutils
    x = rutils.direct(A, b, I=m.interior_nodes())
    print("test_asm.py:163 max(x) =", repr(np.max(x)), "(reference pins 0.073614737354524146)")
    # error norms of nodal interpolants (spfem/assembly.py:1353-1485)
    import json
    norms = {}
    for mname, ename, ref, dim in (("MeshTri", "TriP1", 4, 2), ("MeshTri", "TriP2", 3, 2),
                                   ("MeshTet", "TetP1", 2, 3), ("MeshQuad", "Q1", 3, 2)):
        m = getattr(rmesh, mname)()
        m.refine(ref)
        a = rasm.AssemblerElement(m, ref_element(ename, 1))
        ex, dex = cases.exact_fun(dim)
        N = a.dofnum_u.N
        uh = cases.nodal_values(m, a, ex)
        norms[mname + "_" + ename] = {"L2": float(a.L2error(uh, ex)),
                                      "H1": float(a.H1error(uh, dex))}
        print("norms", mname, ename, norms[mname + "_" + ename])
    # element-wise norms (spfem/assembly.py:1273-1351)
    for mname, ename, ref in cases.INORM_CASES:
        m = getattr(rmesh, mname)()
        m.refine(ref)
        a = rasm.AssemblerElement(m, ref_element(ename, 1))
        N = a.dofnum_u.N
        interp = {0: cases._interp_vec(N, 3)}
        if m.p.shape[0] == 3:
            val = a.inorm(lambda u, du, x, h: cases.inorm_form(u, du, x, h) + du[0][2], interp)
        else:
            val = a.inorm(cases.inorm_form, interp)
        np.savez_compressed(os.path.join(HERE, "inorm_%s_%s.npz" % (mname, ename)), data=val)
        print("inorm", mname, ename, val.shape, float(val.sum()))
    # facet-wise norms (spfem/assembly.py:1185-1271)
    for mname, ename, ref in cases.FNORM_CASES:
        m = getattr(rmesh, mname)()
        m.refine(ref)
        a = rasm.AssemblerElement(m, ref_element(ename, 1))
        interp = {0: cases._interp_vec(a.dofnum_u.N, 3)}
        vb, fb = a.fnorm(cases.fnorm_boundary_form, interp)
        vi, fi = a.fnorm(cases.fnorm_interior_form, interp, interior=True)
        np.savez_compressed(os.path.join(HERE, "fnorm_%s_%s.npz" % (mname, ename)),
                            boundary=vb, boundary_find=fb, interior=vi, interior_find=fi)
        print("fnorm", mname, ename, vb.shape, float(vb.sum()), vi.shape, float(vi.sum()))
    with open(os.path.join(HERE, "error_norms.json"), "w") as fh:
        json.dump(norms, fh, indent=1)


if __name__ == "__main__":
    main()
