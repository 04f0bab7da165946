# -*- coding: utf-8 -*-
"""H3 of SURVEY.md 7 / Appendix B: forms that apply *opaque* callables to the
point data -- things the symbolic tracer cannot follow -- fall back to
host-evaluated coefficient fields (``tracer.FieldContext``): ``x``, ``h``, ``w``
are handed to the form as the reference's own (elements, points) arrays and
whatever it computes from them is streamed to the kernel.  Reference usage:
``spfem/mesh.py:885-901`` (``MeshTri.interpolator``: a matplotlib
``LinearTriInterpolator`` called on ``x``), ``spfem/test_module.py:759-765``
(SymPy ``lambdify``'d load functions)."""
import numpy as np
import pytest

import cases
import hostemu
import util
import asm_oracle
import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement, TraceError


XP = np.linspace(0.0, 1.0, 7)
FP = np.array([0.0, 2.0, -1.0, 0.5, 3.0, 1.0, 0.25])


def load_interp(v, x):                       # np.interp is not a ufunc: untraceable
    return np.interp(x[0], XP, FP) * v


def mass_table(u, v, x):                     # piecewise-constant coefficient by table look-up
    table = np.array([1.0, 10.0, 0.1, 5.0])
    idx = np.minimum((4 * x[0]).astype(int), 3)
    return table[idx] * u * v


def stiff_where(du, dv, x, h):               # np.where + a Python-level branch on data
    k = np.where(x[0] + x[1] > 1.0, 3.0, 1.0)
    if np.any(h > 10):
        k = 0 * k
    return k * (du[0] * dv[0] + du[1] * dv[1])


def load_lambdify(v, x):                     # spfem/test_module.py:759-765
    sympy = pytest.importorskip("sympy")
    X, Y = sympy.symbols("x y")
    uex = (sympy.sin(sympy.pi * X) * sympy.sin(sympy.pi * Y)) ** 2
    load = sympy.Piecewise((sympy.diff(uex, X, 2), X < 0.5), (sympy.diff(uex, Y, 2), True))
    fun = sympy.lambdify((X, Y), load, "numpy")
    return fun(x[0], x[1]) * v


def nonlinear_w(v, w, x):                    # an interpolated previous solution through np.interp
    return np.interp(w[0], XP, FP) * x[0] * v


def vec_interp(u, v, x):
    c = np.interp(x[1], XP, FP)
    return c * (u[0] * v[0] + 2 * u[1] * v[1])


CASES = {
    "load_interp": (load_interp, "TriP1", 1, 3, None),
    "mass_table": (mass_table, "TriP2", 1, 2, None),
    "stiff_where": (stiff_where, "TriP1", 1, 3, None),
    "load_lambdify": (load_lambdify, "TriP2", 1, 2, None),
    "nonlinear_w": (nonlinear_w, "TriP1", 1, 3, "w"),
    "vec_interp": (vec_interp, "TriP1", 2, 2, None),
}


def _setup(name):
    form, ename, ncomp, refine, wmode = CASES[name]
    m = sp.MeshTri()
    m.refine(refine)
    a = AssemblerElement(m, util.element(ename, ncomp))
    interp = None
    if wmode:
        rng = np.random.default_rng(5)
        interp = {0: rng.random(a.dofnum_u.N)}
    ref = asm_oracle.iasm(m, ename, form, ncomp_u=ncomp, ncomp_v=ncomp, interp=interp)
    return a, form, interp, ref


@pytest.mark.parametrize("name", sorted(CASES))
def test_opaque_form_host(name):
    a, form, interp, ref = _setup(name)
    plan = a.plan_iasm(form, interp=interp)
    assert len(plan.fields) >= 1 and plan.spec.nfld == len(plan.fields)
    util.check(hostemu.iasm(a, form, interp=interp), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_opaque_form_cuda(name):
    a, form, interp, ref = _setup(name)
    util.check(a.iasm(form, interp=interp), ref)
    # element subsets go through the same kernels
    if name == "mass_table":
        tind = np.arange(0, a.mesh.t.shape[1], 3)
        util.check(a.iasm(form, tind=tind),
                   asm_oracle.iasm(a.mesh, "TriP2", form, tind=tind))


def test_traceable_forms_do_not_use_fields():
    m = sp.MeshTri()
    m.refine(2)
    a = AssemblerElement(m, sp.ElementTriP1())
    plan = a.plan_iasm(lambda v, x: np.sin(np.pi * x[0]) * (x[1] > 0.5) * v)
    assert plan.fields == [] and plan.spec.nfld == 0


def test_field_mode_keeps_multilinearity_errors():
    m = sp.MeshTri()
    m.refine(1)
    a = AssemblerElement(m, sp.ElementTriP1())
    with pytest.raises(TraceError):
        a.plan_iasm(lambda u, v: abs(u) * v)
    with pytest.raises(TraceError):
        a.plan_iasm(lambda u, v, x: np.interp(x[0], XP, FP) * u * u * v)


def test_same_kernel_for_different_opaque_functions():
    """The generated source depends on the structure only: any opaque load
    ``f(x) * v`` runs the same compiled kernel."""
    m = sp.MeshTri()
    m.refine(2)
    a = AssemblerElement(m, sp.ElementTriP1())
    p1 = a.plan_iasm(load_interp)
    p2 = a.plan_iasm(lambda v, x: np.interp(x[1], XP, FP[::-1]) * v)
    assert p1.source == p2.source
    assert not np.array_equal(p1.fields[0], p2.fields[0])


# ---------------------------------------------------------------------------
# facet forms (fasm): opaque functions of x, h, n
# ---------------------------------------------------------------------------
def bnd_interp(u, v, x, h):
    return np.interp(x[1], XP, FP) / h * u * v


def bnd_load_table(v, x, n):
    g = np.where(x[0] > 0.5, 2.0, -1.0) * n[0] + np.interp(x[1], XP, FP) * n[1]
    return g * v


def jump_opaque(u1, u2, v1, v2, x, h):
    # (not a pure jump: on a continuous element (u1 - u2)(v1 - v2) assembles to round-off)
    k = np.interp(x[0] + x[1], [0.0, 1.0, 2.0], [1.0, 4.0, 2.0])
    return k / h * (u1 * v1 + 2.0 * u2 * v2) - k * u1 * v2


FCASES = {
    "bnd_interp_p2": (bnd_interp, "MeshTri", 2, "TriP2", False),
    "bnd_load_table_p1": (bnd_load_table, "MeshTri", 3, "TriP1", False),
    "bnd_interp_tet": (bnd_interp, "MeshTet", 1, "TetP1", False),
    "jump_opaque_p1": (jump_opaque, "MeshTri", 2, "TriP1", True),
}


def _fsetup(name):
    form, mesh, refine, ename, interior = FCASES[name]
    m = getattr(sp, mesh)()
    m.refine(refine)
    a = AssemblerElement(m, util.element(ename, 1))
    ref = asm_oracle.fasm(m, ename, form, interior=interior)
    return a, form, interior, ref


@pytest.mark.parametrize("name", sorted(FCASES))
def test_opaque_facet_form_host(name):
    a, form, interior, ref = _fsetup(name)
    plan = a.plan_fasm(form, interior=interior)
    assert len(plan.fields) >= 1
    util.check(hostemu.fasm(a, form, interior=interior), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FCASES))
def test_opaque_facet_form_cuda(name):
    a, form, interior, ref = _fsetup(name)
    util.check(a.fasm(form, interior=interior), ref)
    if name == "bnd_interp_p2":          # a facet subset gets its own fields
        find = a.mesh.boundary_facets()[::2]
        util.check(a.fasm(form, find=find), asm_oracle.fasm(a.mesh, "TriP2", form, find=find))
