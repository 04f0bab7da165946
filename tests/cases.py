# -*- coding: utf-8 -*-
"""Shared parity cases: the same (mesh, element pair, form, options) tuples are
run through the patched reference (``tests/golden/make_golden.py`` -> golden
``.npz``), the oracle, the host emulation of the generated kernels and the
CUDA path.  Forms are the idioms of the reference's own tests
(SURVEY.md appendix B): spfem/test_asm.py, test_mapping.py, test_module.py,
README.md and learning/Example 1."""
import numpy as np

GAMMA = 100.0


def lap2(du, dv):
    return du[0] * dv[0] + du[1] * dv[1]


def lap3(du, dv):
    return du[0] * dv[0] + du[1] * dv[1] + du[2] * dv[2]


def mass(u, v):
    return u * v


def rt0_sigtau(u, v):
    return u[0] * v[0] + u[1] * v[1]


def rt0_divsigv(du, v):
    return du * v


def load1(v):
    return 1.0 * v


def load_sin_h(v, x, h):
    return h * np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]) * v


def mass_x2(u, v, x):
    return x[0] ** 2 * u * v


def load_masks(v, x):
    # boolean masks as in spfem/test_asm.py:200-201, test_module.py:284-288
    return ((x[0] >= 0.4) & (x[0] <= 0.6) & (x[1] == 1.0)) * v + (x[0] == 1) * 1.0 * v


def mass_w(u, v, w):
    return w[0] * u * v


def load_w2(v, w):
    return w[7] * w[7] * v


def _eps(dw):
    return [[dw[0][0], 0.5 * (dw[0][1] + dw[1][0])],
            [0.5 * (dw[0][1] + dw[1][0]), dw[1][1]]]


def elasticity(du, dv):
    # 2 mu eps(u):eps(v) + lambda tr eps(u) tr eps(v), lambda = mu = 1
    eu, ev = _eps(du), _eps(dv)
    return (2.0 * sum(eu[i][j] * ev[i][j] for i in range(2) for j in range(2))
            + 1.0 * (eu[0][0] + eu[1][1]) * (ev[0][0] + ev[1][1]))


def stokes_b(du, v):
    return (du[0][0] + du[1][1]) * v


def vec3_mixed(u, v, du, dv):
    return (u[0] * v[0] + u[1] * v[1] + u[2] * v[2] + du[0][1] * dv[1][0]
            + du[2][2] * dv[0][0])


def q_load(v, x, h):
    return h * np.exp(x[0]) * x[1] * v


def bnd_mass(u, v):
    return u * v


def nitsche(u, v, du, dv, x, h, n):
    # spfem/test_module.py:498-501
    return (GAMMA * 1 / h * u * v - du[0] * n[0] * v - du[1] * n[1] * v
            - u * dv[0] * n[0] - u * dv[1] * n[1])


def bnd_lin(v, dv, x, h, n):
    return GAMMA / h * np.sin(x[0]) * v - dv[0] * n[0] * x[1]


def normal0(v, n):
    return n[0] * v


def normal2(v, n):
    return n[2] * v


def sipg(u1, u2, v1, v2, du1, du2, dv1, dv2, h, n):
    # spfem/test_asm.py:38-44 style interior penalty terms
    # (signs chosen so that the terms do not cancel for continuous elements)
    return (10 / h * (u1 + 2 * u2) * (v1 - 3 * v2)
            - 0.5 * ((du1[0] + du2[0]) * n[0] + (du1[1] + du2[1]) * n[1]) * (v1 + v2))


def facet_w(v, w, dw, n):
    return w[0] * v + (dw[0][0] * n[0] + dw[0][1] * n[1]) * v


def vec_nitsche(u, v, du, dv, h, n):
    return (GAMMA / h * (u[0] * v[0] + u[1] * v[1])
            - sum(du[i][j] * n[j] * v[i] for i in range(2) for j in range(2))
            - sum(dv[i][j] * n[j] * u[i] for i in range(2) for j in range(2)))


def tet_bnd(u, v, du, dv, x, h, n):
    return (1 / h * u * v - (du[0] * n[0] + du[1] * n[1] + du[2] * n[2]) * v
            + x[2] * u * v)


def jump(u1, u2, v1, v2, h):
    return 1 / h * (u1 + 2 * u2) * (v1 - 3 * v2)


def q_bnd(u, v, x, h):
    return 1 / h * u * v * x[0]


def q_bnd_lin(v, dv, x):
    return x[1] * v + dv[0]


def _interp_vec(n, seed=0):
    return np.random.default_rng(seed).standard_normal(n)


class Case(object):
    def __init__(self, name, mesh, refine, eu, form, kind="iasm", ncu=1,
                 ev=None, ncv=1, interp_key=None, tind_step=None,
                 find_step=None, oracle=True, **kw):
        self.oracle = oracle      # False: element not in the oracle, golden only
        self.name, self.mesh, self.refine = name, mesh, refine
        self.eu, self.ev, self.ncu, self.ncv = eu, ev, ncu, ncv
        self.form, self.kind, self.kw = form, kind, kw
        self.interp_key, self.tind_step, self.find_step = interp_key, tind_step, find_step


CASES = [
    Case("trip1_lap", "MeshTri", 4, "TriP1", lap2),
    Case("trip1_mass", "MeshTri", 4, "TriP1", mass),
    Case("trip1_lap_r6", "MeshTri", 6, "TriP1", lap2),
    Case("trip2_lap", "MeshTri", 3, "TriP2", lap2),
    Case("trip2_mass", "MeshTri", 3, "TriP2", mass),
    Case("tetp1_lap", "MeshTet", 2, "TetP1", lap3),
    Case("tetp2_lap", "MeshTet", 2, "TetP2", lap3),
    Case("tetp2_mass", "MeshTet", 2, "TetP2", mass),
    Case("q1_lap", "MeshQuad", 3, "Q1", lap2),
    Case("q2_lap", "MeshQuad", 3, "Q2", lap2),
    Case("q2_mass", "MeshQuad", 3, "Q2", mass),
    Case("trip1_load", "MeshTri", 4, "TriP1", load1),
    Case("trip1_load_sin_h", "MeshTri", 4, "TriP1", load_sin_h),
    Case("trip1_mass_x2", "MeshTri", 4, "TriP1", mass_x2),
    Case("trip1_masks", "MeshTri", 4, "TriP1", load_masks),
    Case("trip1_interp_mass", "MeshTri", 4, "TriP1", mass_w, interp_key=0),
    Case("trip1_interp_load", "MeshTri", 4, "TriP1", load_w2, interp_key=7),
    Case("trip1_tind", "MeshTri", 4, "TriP1", lap2, tind_step=3),
    Case("trip1_lap_order5", "MeshTri", 3, "TriP1", lap2, intorder=5),
    Case("vecp2_elasticity", "MeshTri", 3, "TriP2", elasticity, ncu=2, ncv=2),
    Case("stokes_b", "MeshTri", 3, "TriP2", stokes_b, ncu=2, ev="TriP1", ncv=1),
    Case("vectetp1_mixed", "MeshTet", 2, "TetP1", vec3_mixed, ncu=3, ncv=3),
    Case("q1_load", "MeshQuad", 3, "Q1", q_load),
    # facets
    Case("f_trip1_mass", "MeshTri", 3, "TriP1", bnd_mass, kind="fasm"),
    Case("f_trip1_nitsche", "MeshTri", 3, "TriP1", nitsche, kind="fasm"),
    Case("f_trip1_lin", "MeshTri", 3, "TriP1", bnd_lin, kind="fasm"),
    Case("f_trip1_n0", "MeshTri", 3, "TriP1", normal0, kind="fasm"),
    Case("f_trip1_sipg", "MeshTri", 3, "TriP1", sipg, kind="fasm", interior=True),
    Case("f_trip1_subset", "MeshTri", 3, "TriP1", bnd_mass, kind="fasm", find_step=2),
    Case("f_trip1_interp", "MeshTri", 3, "TriP1", facet_w, kind="fasm", interp_key=0),
    Case("f_trip2_nitsche", "MeshTri", 3, "TriP2", nitsche, kind="fasm"),
    Case("f_vecp2_nitsche", "MeshTri", 2, "TriP2", vec_nitsche, kind="fasm", ncu=2, ncv=2),
    Case("f_tetp1_n2", "MeshTet", 2, "TetP1", normal2, kind="fasm"),
    Case("f_tetp1_bnd", "MeshTet", 2, "TetP1", tet_bnd, kind="fasm"),
    Case("f_tetp2_bnd", "MeshTet", 1, "TetP2", tet_bnd, kind="fasm"),
    Case("f_tetp2_jump", "MeshTet", 1, "TetP2", jump, kind="fasm", interior=True),
    Case("f_q1_bnd", "MeshQuad", 3, "Q1", q_bnd, kind="fasm", normals=False),
    Case("f_q2_lin", "MeshQuad", 2, "Q2", q_bnd_lin, kind="fasm", normals=False),
    # elements outside the oracle's six: discontinuous P1, MINI, P0 (golden only)
    Case("tridg1_lap", "MeshTri", 2, "TriDG1", lap2, oracle=False),
    Case("f_tridg1_sipg", "MeshTri", 2, "TriDG1", sipg, kind="fasm", interior=True, oracle=False),
    Case("trimini_lap", "MeshTri", 2, "TriMini", lap2, oracle=False),
    Case("trimini_mass", "MeshTri", 2, "TriMini", mass, oracle=False),
    Case("trip0_mass", "MeshTri", 2, "TriP0", mass, oracle=False),
    Case("stokes_b_p0", "MeshTri", 2, "TriP2", stokes_b, ncu=2, ev="TriP0", ncv=1, oracle=False),
    Case("tetp0_mass", "MeshTet", 1, "TetP0", mass, oracle=False),
    # Raviart-Thomas (Piola map; the mixed Poisson blocks of test_module.py:18-77)
    Case("rt0_sigtau", "MeshTri", 2, "TriRT0", rt0_sigtau, oracle=False),
    Case("rt0_div_p0", "MeshTri", 2, "TriRT0", rt0_divsigv, ev="P0", oracle=False),
    # hierarchical p-element (spfem/element.py:623-764)
    Case("tripp3_lap", "MeshTri", 2, "TriPp3", lap2, oracle=False),
    Case("tripp4_mass", "MeshTri", 1, "TriPp4", mass, oracle=False),
    Case("tripp3_load", "MeshTri", 2, "TriPp3", load_sin_h, oracle=False),
]
BY_NAME = {c.name: c for c in CASES}


def case_inputs(case, mesh, ndofs_u):
    """Keyword arguments (interp / tind / find) of a case for a given mesh."""
    kw = dict(case.kw)
    if case.interp_key is not None:
        kw["interp"] = {case.interp_key: _interp_vec(ndofs_u)}
    if case.tind_step:
        kw["tind"] = np.arange(0, mesh.t.shape[1], case.tind_step)
    if case.find_step:
        bnd = np.nonzero(mesh.f2t[1] == -1)[0]
        kw["find"] = bnd[::case.find_step]
    return kw


def compare_csr(A, B, rtol=1e-12):
    """Parity metric of SURVEY.md 8(c): identical pattern and index width; per
    stored entry ``|a-b| <= rtol * max(|a|, |b|, s_row)`` with ``s_row`` the
    largest magnitude in the oracle row.  Only rows that are *entirely*
    round-off (``s_row < 1e-10 * max|B|``: e.g. the jump terms of a continuous
    element, exact zeros up to cancellation noise in both implementations)
    are compared absolutely, against ``1e-3 * rtol * max|B|`` -- a relative test
    of noise against noise is meaningless."""
    assert A.shape == B.shape, (A.shape, B.shape)
    assert A.indptr.dtype == B.indptr.dtype and A.indices.dtype == B.indices.dtype, \
        (A.indptr.dtype, B.indptr.dtype)
    assert np.array_equal(A.indptr, B.indptr), "indptr differs"
    assert np.array_equal(A.indices, B.indices), "indices differ"
    if A.nnz == 0:
        return 0.0
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    srow = np.zeros(A.shape[0])
    np.maximum.at(srow, rows, np.abs(B.data))
    gmax = np.abs(B.data).max()
    scale = np.maximum(np.maximum(np.abs(A.data), np.abs(B.data)), srow[rows])
    noise = srow[rows] < 1e-10 * gmax
    scale = np.where(noise, 1e-3 * gmax, scale)
    err = np.abs(A.data - B.data) / np.maximum(scale, 1e-300)
    assert err.max() <= rtol, "max relative error %.3e" % err.max()
    return float(err.max())


def compare_vec(a, b, rtol=1e-12):
    assert a.shape == b.shape
    err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
    assert err <= rtol, "max relative error %.3e" % err
    return float(err)


def exact_fun(dim):
    """Smooth exact solution and its gradient (callables of the ``x`` dict)."""
    if dim == 2:
        ex = lambda x: np.sin(np.pi * x[0]) * np.cos(0.5 * x[1]) + x[0] * x[1]
        dex = {0: lambda x: np.pi * np.cos(np.pi * x[0]) * np.cos(0.5 * x[1]) + x[1],
               1: lambda x: -0.5 * np.sin(np.pi * x[0]) * np.sin(0.5 * x[1]) + x[0]}
    else:
        ex = lambda x: np.sin(x[0]) * x[1] + np.exp(0.3 * x[2])
        dex = {0: lambda x: np.cos(x[0]) * x[1], 1: lambda x: np.sin(x[0]) + 0 * x[1],
               2: lambda x: 0.3 * np.exp(0.3 * x[2])}
    return ex, dex


def nodal_values(mesh, asm, ex):
    """A discrete function: the exact solution at the vertices, mid-entity
    averages for the remaining DOFs (any fixed vector works for the test)."""
    N = asm.dofnum_u.N
    uh = np.zeros(N)
    nv = mesh.p.shape[1]
    uh[:nv] = ex({k: mesh.p[k] for k in range(mesh.p.shape[0])})
    if N > nv:
        uh[nv:] = 0.5 * np.cos(np.arange(N - nv))
    return uh


def inorm_form(u, du, x, h):
    # residual-estimator like: h * (f + something of the discrete solution).
    # One interpolated field only: with several keys the reference aliases all
    # ``w[k]`` to one array (``w[k] = zero`` followed by in-place ``+=``,
    # spfem/assembly.py:1334-1346), which is not reproduced here.
    return h * (np.sin(x[0]) + u[0] * x[1]) + du[0][0] - 2.0 * du[0][1]


def fnorm_boundary_form(u, du, x, n, t, h):
    # Neumann-residual like (a posteriori estimators): normal / tangential
    # derivatives of the interpolated solution on boundary facets
    val = h * (u[0] * x[0] + sum(du[0][k] * n[k] for k in range(len(x))))
    if len(t) > 0:
        val = val + du[0][0] * t[0] + du[0][1] * t[1]
    return val


def fnorm_interior_form(u1, u2, du1, du2, x, n, t, h):
    # jump of the normal derivative across interior facets
    jump = sum((du1[0][k] - du2[0][k]) * n[k] for k in range(len(x)))
    return h * jump + u1[0] - 0.5 * u2[0] + x[0]


FNORM_CASES = [("MeshTri", "TriP1", 3), ("MeshTri", "TriP2", 2), ("MeshTet", "TetP1", 1),
               ("MeshTet", "TetP2", 1)]
INORM_CASES = [("MeshTri", "TriP1", 3), ("MeshTri", "TriP2", 2), ("MeshTet", "TetP2", 1),
               ("MeshQuad", "Q2", 2)]
