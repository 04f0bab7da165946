# -*- coding: utf-8 -*-
"""Quadrilateral elements (``MappingQ1``, spfem/mapping.py:47-191): the interior
kernel is a loop over the quadrature points with the reference tables in
constant / global memory (``codegen.looped_quad_body``).  Forms with every kind
of point data against the oracle, on the host emulation and on the GPU, and the
straight-line generator (``SPFEM_QLOOP=0``) as a second opinion."""
import numpy as np
import pytest

import cases
import hostemu
import util
import asm_oracle
import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement


def f_x(u, v, du, dv, x, h):
    return (1.0 + x[0] * x[1]) * (du[0] * dv[0] + du[1] * dv[1]) + np.sin(x[0]) * h * u * v


def f_w(u, v, w):
    return (1.0 + w[0] * w[0]) * u * v


def f_load(v, dv, x):
    return np.cos(x[0] + x[1]) * v + x[0] * dv[1]


def f_vec(du, dv):
    e = lambda d: [[d[0][0], 0.5 * (d[0][1] + d[1][0])], [0.5 * (d[0][1] + d[1][0]), d[1][1]]]
    eu, ev = e(du), e(dv)
    return 2.0 * sum(eu[i][j] * ev[i][j] for i in range(2) for j in range(2)) + \
        (eu[0][0] + eu[1][1]) * (ev[0][0] + ev[1][1])


def f_nonsym(u, dv, x):
    return x[1] * u * dv[0]


def f_opaque(u, v, x):
    return np.interp(x[0], [0.0, 0.5, 1.0], [1.0, 3.0, 2.0]) * u * v


CASES = {
    "q1_x": (f_x, "Q1", 1, False), "q2_x": (f_x, "Q2", 1, False),
    "q1_w": (f_w, "Q1", 1, True), "q2_w": (f_w, "Q2", 1, True),
    "q2_load": (f_load, "Q2", 1, False),
    "vecq1_elasticity": (f_vec, "Q1", 2, False),
    "q2_nonsym": (f_nonsym, "Q2", 1, False),
    "q1_opaque": (f_opaque, "Q1", 1, False),
    "q2_laplace": (cases.lap2, "Q2", 1, False),
    "q2_convection": (lambda u, dv, h: (0.3 + h) * u * dv[0] - 2.0 * u * dv[1], "Q2", 1, False),
}


def _setup(name, jiggle=True):
    form, ename, ncomp, use_w = CASES[name]
    m = sp.MeshQuad()
    m.refine(2)
    if jiggle:      # genuinely bilinear maps: the Jacobian varies from point to point
        I = m.interior_nodes()
        m.p[:, I] += 0.03 * np.random.default_rng(2).standard_normal((2, len(I)))
    a = AssemblerElement(m, util.element(ename, ncomp))
    interp = {0: np.random.default_rng(3).random(a.dofnum_u.N)} if use_w else None
    ref = asm_oracle.iasm(m, ename, form, ncomp_u=ncomp, ncomp_v=ncomp, interp=interp)
    return a, form, interp, ref


@pytest.mark.parametrize("name", sorted(CASES))
def test_quadrature_loop_host(name, monkeypatch):
    a, form, interp, ref = _setup(name)
    plan = a.plan_iasm(form, interp=interp)
    assert plan.spec.looped and "for (int q = 0;" in plan.source
    util.check(hostemu.iasm(a, form, interp=interp), ref)
    monkeypatch.setenv("SPFEM_QLOOP", "0")
    plan0 = a.plan_iasm(form, interp=interp)
    assert not plan0.spec.looped
    util.check(hostemu.iasm(a, form, interp=interp), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_quadrature_loop_cuda(name):
    a, form, interp, ref = _setup(name)
    util.check(a.iasm(form, interp=interp), ref)                      # tile kernel
    tind = np.arange(a.mesh.t.shape[1])
    util.check(a.iasm(form, interp=interp, tind=tind), ref)           # two-kernel route


@pytest.mark.parametrize("name", ["q2_laplace", "q2_convection", "vecq1_elasticity", "q2_x"])
def test_parallelogram_shortcut_host(name):
    """Uniform grid: every element is a parallelogram, the Jacobian and the
    coefficients are evaluated at the first point only (forms without x / w)."""
    a, form, interp, ref = _setup(name, jiggle=False)
    plan = a.plan_iasm(form, interp=interp)
    assert ("spfem_affine" in plan.source) == (name != "q2_x")
    assert bool(plan.spec.affine_variant) == (name != "q2_x")
    util.check(hostemu.iasm(a, form, interp=interp), ref)
    if plan.spec.affine_variant:
        # the kernel variant for meshes of parallelograms only (no loop branch compiled)
        import copy
        pv = copy.copy(plan)
        pv.source = "#define SPFEM_PARALLELOGRAMS 1\n" + plan.source
        assert np.array_equal(hostemu.run_interior(a, pv), hostemu.run_interior(a, plan))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["q2_laplace", "q2_convection", "vecq1_elasticity"])
def test_parallelogram_shortcut_cuda(name):
    a, form, interp, ref = _setup(name, jiggle=False)
    util.check(a.iasm(form, interp=interp), ref)
    # a mesh that mixes parallelograms and general quadrilaterals in one warp
    m = sp.MeshQuad()
    m.refine(3)
    I = m.interior_nodes()[::3]
    m.p[:, I] += 0.01 * np.random.default_rng(4).standard_normal((2, len(I)))
    ename, ncomp = CASES[name][1], CASES[name][2]
    a2 = AssemblerElement(m, util.element(ename, ncomp))
    util.check(a2.iasm(form), asm_oracle.iasm(m, ename, form, ncomp_u=ncomp, ncomp_v=ncomp))


@pytest.mark.gpu
def test_parallelogram_variant_follows_the_mesh():
    """Uniform mesh: the engine launches the variant compiled without the quadrature
    loop; after the coordinates change (``upload_mesh``) the general kernel."""
    m = sp.MeshQuad()
    m.refine(3)
    a = AssemblerElement(m, sp.ElementQ2())
    prep = a.prepare_iasm(cases.lap2)
    eng = a._get_engine()
    assert type(prep).__name__ == "TiledPrepared" and eng.all_parallelograms()
    k0 = prep.tile_kernel.value
    prep.launch()
    util.check(prep.result(), asm_oracle.iasm(m, "Q2", cases.lap2))
    I = m.interior_nodes()[::3]
    m.p[:, I] += 0.01 * np.random.default_rng(5).standard_normal((2, len(I)))
    eng.upload_mesh()
    assert not eng.all_parallelograms()
    prep.launch()
    assert prep.tile_kernel.value != k0
    util.check(prep.result(), asm_oracle.iasm(m, "Q2", cases.lap2))
    util.check(a.iasm(cases.lap2), asm_oracle.iasm(m, "Q2", cases.lap2))


def test_symmetric_entries_are_aliased_in_the_loop():
    a, form, interp, ref = _setup("q2_x")
    spec = a.plan_iasm(form).spec
    assert spec.n_unique == 45 and len(set(spec.alias)) == 45          # 81 entries, 45 distinct
    spec = a.plan_iasm(f_nonsym).spec
    assert spec.n_unique == 81


# ---------------------------------------------------------------------------
# randomly composed geometry-only forms (constants and the element size h) on
# uniform, sheared (parallelograms that are not rectangles) and mixed meshes:
# the loop-free parallelogram branch against the oracle
# ---------------------------------------------------------------------------
def _geom_form(rng, bilinear):
    terms = []
    for _ in range(int(rng.integers(1, 4))):
        a = float(np.round(rng.uniform(0.5, 2.0), 3))
        kind = int(rng.integers(3))
        terms.append((a, kind, int(rng.integers(3)), int(rng.integers(3))))

    def coef(a, kind, h):
        return a if kind == 0 else (a / h if kind == 1 else a + h * h)

    def pick(val, grad, t):
        return val if t == 0 else grad[t - 1]
    if bilinear:
        def form(u, v, du, dv, h):
            return sum(coef(a, k, h) * pick(u, du, tu) * pick(v, dv, tv) for a, k, tu, tv in terms)
    else:
        def form(v, dv, h):
            return sum(coef(a, k, h) * pick(v, dv, tv) for a, k, tu, tv in terms)
    return form


def _geom_mesh(kind):
    m = sp.MeshQuad()
    m.refine(2)
    if kind in ("sheared", "mixed"):
        m.p[0] += 0.25 * m.p[1]                    # dyadic shear: still exact parallelograms
        m.p[1] *= 1.5
    if kind == "mixed":
        I = m.interior_nodes()[::4]
        m.p[:, I] += 0.02 * np.random.default_rng(9).standard_normal((2, len(I)))
    return m


def _geom_cases():
    rng = np.random.default_rng(777)
    return [(e, mk, k, bool(k % 2 == 0), int(rng.integers(1 << 30)))
            for e in ("Q1", "Q2") for mk in ("uniform", "sheared", "mixed") for k in range(2)]


def _geom_run(ename, mkind, bilinear, seed, device):
    m = _geom_mesh(mkind)
    form = _geom_form(np.random.default_rng(seed), bilinear)
    a = AssemblerElement(m, util.element(ename, 1))
    plan = a.plan_iasm(form)
    assert plan.spec.affine_variant
    ref = asm_oracle.iasm(m, ename, form)
    got = a.iasm(form) if device else hostemu.iasm(a, form)
    if bilinear:
        cases.compare_csr(got, ref)
    else:
        cases.compare_vec(got, ref)
    if device:
        assert a._get_engine().all_parallelograms() == (mkind != "mixed")


@pytest.mark.parametrize("ename,mkind,k,bilinear,seed", _geom_cases())
def test_random_geometry_only_form_host(ename, mkind, k, bilinear, seed):
    _geom_run(ename, mkind, bilinear, seed, False)


@pytest.mark.gpu
@pytest.mark.parametrize("ename,mkind,k,bilinear,seed", _geom_cases())
def test_random_geometry_only_form_cuda(ename, mkind, k, bilinear, seed):
    _geom_run(ename, mkind, bilinear, seed, True)
