# -*- coding: utf-8 -*-
"""The reference's own known-answer tests for the assembly path
(spfem/test_asm.py:143-310, spfem/test_mapping.py:22-64), restated against this
package: once with the generated kernels compiled for the host (CPU suite) and
once through the public API on the GPU (``-m gpu``)."""
import numpy as np
import pytest
import scipy.sparse.linalg

import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
import hostemu


class _Host(object):
    """iasm / fasm of an assembler through the host emulation."""

    def __init__(self, a):
        self.a = a

    def iasm(self, form, **kw):
        return hostemu.iasm(self.a, form, **kw)

    def fasm(self, form, **kw):
        return hostemu.fasm(self.a, form, **kw)


def _asm(mesh, elem, device):
    a = AssemblerElement(mesh, elem)
    return a if device else _Host(a)


def _unit_square(refine):
    m = sp.MeshTri()
    m.refine(refine)
    p = m.p
    D = np.nonzero((p[0] == 0) | (p[1] == 0) | (p[0] == 1) | (p[1] == 1))[0]
    I = np.setdiff1d(np.arange(p.shape[1]), D)
    return m, D, I


def _poisson(device):                       # test_asm.py:145-163
    m, D, I = _unit_square(5)
    a = _asm(m, sp.ElementTriP1(), device)
    A = a.iasm(lambda u, v, du, dv, x, h: du[0] * dv[0] + du[1] * dv[1])
    f = a.iasm(lambda v, dv, x, h: 1 * v)
    x = np.zeros(A.shape[0])
    x[I] = scipy.sparse.linalg.spsolve(A[np.ix_(I, I)].tocsc(), f[I])
    assert abs(np.max(x) - 0.073614737354524146) < 1e-7


def _analytic_xy(device):                   # test_asm.py:165-194
    m, D, I = _unit_square(5)
    a = _asm(m, sp.ElementTriP1(), device)
    K = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
    f = a.iasm(lambda v, x: 2 * np.pi ** 2 * np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]) * v)
    x = np.zeros(K.shape[0])
    x[I] = scipy.sparse.linalg.spsolve(K[np.ix_(I, I)].tocsc(), f[I])
    assert abs(np.max(x - np.sin(np.pi * m.p[0]) * np.sin(np.pi * m.p[1]))) < 5e-4


def _full_poisson(device):                  # test_asm.py:196-224
    m, _, _ = _unit_square(5)
    F = lambda x, y: 100.0 * ((x >= 0.4) & (x <= 0.6) & (y >= 0.4) & (y <= 0.6))
    G = lambda x, y: (y == 0) * 1.0 + (y == 1) * (-1.0)
    a = _asm(m, sp.ElementTriP1(), device)
    K = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
    B = a.fasm(lambda u, v: u * v)
    f = a.iasm(lambda v, x: F(x[0], x[1]) * v)
    g = a.fasm(lambda v, x: G(x[0], x[1]) * v)
    D = np.nonzero(m.p[0, :] == 0)[0]
    I = np.setdiff1d(np.arange(0, m.p.shape[1]), D)
    x = np.zeros(K.shape[0])
    x[I] = scipy.sparse.linalg.spsolve((K[np.ix_(I, I)] + B[np.ix_(I, I)]).tocsc(), f[I] + g[I])
    assert abs(np.max(x) - 1.89635971369) < 5e-3


def _interp(device):                        # test_asm.py:226-269
    m, _, _ = _unit_square(5)
    a = _asm(m, sp.ElementTriP1(), device)
    u = np.sin(np.pi * m.p[0, :])
    f = a.iasm(lambda v, w: w[0] * v, interp={0: u})
    A = a.iasm(lambda u, v: u * v)
    assert np.linalg.norm(f - A * u) < 1e-10
    m = sp.MeshTri()
    m.refine(3)
    a = _asm(m, sp.ElementTriP1(), device)
    u = np.sin(np.pi * m.p[0, :])
    f = a.fasm(lambda v, w: w[0] * v, interp={0: u})
    A = a.fasm(lambda u, v: u * v)
    assert np.linalg.norm(f - A * u) < 1e-10


def _subset(device):                        # test_asm.py:272-310
    m = sp.MeshTri()
    m.refine(4)
    I1 = np.arange(m.t.shape[1] // 2)
    I2 = np.setdiff1d(np.arange(m.t.shape[1]), I1)
    bix = m.boundary_facets()
    bix1 = bix[0:len(bix) // 2]
    bix2 = np.setdiff1d(bix, bix1)
    a = _asm(m, sp.ElementTriP1(), device)
    dudv = lambda du, dv: du[0] * dv[0] + du[1] * dv[1]
    A, A1, A2 = a.iasm(dudv), a.iasm(dudv, tind=I1), a.iasm(dudv, tind=I2)
    B, B1, B2 = a.fasm(dudv), a.fasm(dudv, find=bix1), a.fasm(dudv, find=bix2)
    f = a.iasm(lambda v: 1 * v)
    I = m.interior_nodes()
    x = np.zeros(A.shape[0])
    x[I] = scipy.sparse.linalg.spsolve(((A + B)[I].T[I].T).tocsc(), f[I])
    X = np.zeros(A.shape[0])
    X[I] = scipy.sparse.linalg.spsolve(((A1 + B1)[I].T[I].T + (A2 + B2)[I].T[I].T).tocsc(), f[I])
    assert np.linalg.norm(x - X) < 1e-10


def _normals(device):                       # test_mapping.py:22-64
    m = sp.MeshTri()
    m.refine(2)
    a = _asm(m, sp.ElementTriP1(), device)
    N1 = a.fasm(lambda v, n: n[0] * v, normals=True)
    N2 = a.fasm(lambda v, n: n[1] * v, normals=True)
    one, zero = np.ones(N1.shape[0]), np.zeros(N1.shape[0])
    assert abs(N1.dot(one) + N2.dot(zero)) < 1e-7
    assert abs(N1.dot(zero) + N2.dot(one)) < 1e-7
    m = sp.MeshTet()
    m.refine(2)
    a = _asm(m, sp.ElementTetP1(), device)
    def comp(k):
        return lambda v, n: n[k] * v
    N = [a.fasm(comp(k), normals=True) for k in range(3)]
    one = np.ones(N[0].shape[0])
    assert abs(sum(Nk.dot(one) for Nk in N)) < 1e-7
    for k in range(3):
        assert (N[k][m.p[k, :] == 1.0] >= 0).all()
        assert (N[k][m.p[k, :] == 0.0] <= 0).all()


CHECKS = [_poisson, _analytic_xy, _full_poisson, _interp, _subset, _normals]


@pytest.mark.parametrize("check", CHECKS, ids=lambda f: f.__name__.strip("_"))
def test_reference_known_answer_host(check):
    check(False)


@pytest.mark.gpu
@pytest.mark.parametrize("check", CHECKS, ids=lambda f: f.__name__.strip("_"))
def test_reference_known_answer_cuda(check):
    check(True)


def _h_convergence(elem):
    """The Robin problem of spfem/test_module.py:258-395 (TriP2Test, TriPpTest):
    H1 error of the solution of ``(grad u, grad v) + <u, v> = (f, v) + <g, v>``
    over four refinements; returns the fitted convergence rate."""
    from scipy.sparse.linalg import spsolve
    U = lambda x: 1 + x[0] - x[0] ** 2 * x[1] ** 2
    dexact = {0: lambda x: 1 - 2 * x[0] * x[1] ** 2, 1: lambda x: -2 * x[0] ** 2 * x[1]}
    F = lambda x, y: 2 * x ** 2 + 2 * y ** 2
    G = lambda x, y: ((x == 1) * (3 - 3 * y ** 2) + (x == 0) * 0 + (y == 1) * (1 + x - 3 * x ** 2)
                      + (y == 0) * (1 + x))
    mesh = sp.MeshTri()
    mesh.refine(1)
    hs, errs = [], []
    for _ in range(4):
        mesh.refine()
        a = AssemblerElement(mesh, elem)
        A = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
        f = a.iasm(lambda v, x: F(x[0], x[1]) * v)
        B = a.fasm(lambda u, v: u * v)
        g = a.fasm(lambda v, x: G(x[0], x[1]) * v)
        u = spsolve((A + B).tocsc(), f + g)
        hs.append(mesh.param())
        errs.append(a.H1error(u, dexact))
    return np.polyfit(np.log10(hs), np.log10(errs), 1)[0]


@pytest.mark.gpu
def test_convergence_trip2():               # test_module.py:258-323
    assert _h_convergence(sp.ElementTriP2()) > 0.95 * 2


@pytest.mark.gpu
@pytest.mark.parametrize("p", [1, 2, 3])    # test_module.py:325-395 (p = 1, 2 there)
def test_convergence_tripp(p):
    assert _h_convergence(sp.element.ElementTriPp(p)) >= 0.95 * p


@pytest.mark.gpu
def test_convergence_nitsche():             # test_module.py:475-532
    """Poisson with Nitsche's method for the Dirichlet condition: boundary forms
    with normals, the facet size ``h`` and interior gradients."""
    from scipy.sparse.linalg import spsolve
    G = lambda x, y: x ** 3 * np.sin(20.0 * y)
    U = lambda x: G(x[0], x[1])
    dU = {0: lambda x: 3.0 * x[0] ** 2 * np.sin(20.0 * x[1]),
          1: lambda x: 20.0 * x[0] ** 3 * np.cos(20.0 * x[1])}
    gamma = 100

    def uv(u, v, du, dv, x, h, n):
        return (gamma * 1 / h * u * v - du[0] * n[0] * v - du[1] * n[1] * v
                - u * dv[0] * n[0] - u * dv[1] * n[1])

    def fv(v, dv, x):
        return (-6.0 * x[0] * np.sin(20.0 * x[1]) + 400.0 * x[0] ** 3 * np.sin(20.0 * x[1])) * v

    def gv(v, dv, x, h, n):
        return (G(x[0], x[1]) * v + gamma * 1 / h * G(x[0], x[1]) * v
                - dv[0] * n[0] * G(x[0], x[1]) - dv[1] * n[1] * G(x[0], x[1]))
    mesh = sp.MeshTri()
    mesh.refine(2)
    hs, errs = [], []
    for _ in range(4):
        mesh.refine()
        a = AssemblerElement(mesh, sp.ElementTriP1())
        K = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
        B = a.fasm(uv, normals=True)
        f = a.iasm(fv)
        g = a.fasm(gv, normals=True)
        x = spsolve((K + B).tocsc(), f + g)
        hs.append(mesh.param())
        errs.append(np.sqrt(a.L2error(x, U) ** 2 + a.H1error(x, dU) ** 2))
    assert np.polyfit(np.log10(hs), np.log10(errs), 1)[0] >= 0.99


def _robin_rate(meshes, make_elem, U, dexact, F, G, normals=True):
    """Energy-norm convergence rate of the Robin problem over ``meshes``."""
    from scipy.sparse.linalg import spsolve
    dim = len(dexact)
    hs, errs = [], []
    for mesh in meshes:
        a = AssemblerElement(mesh, make_elem())
        A = a.iasm(lambda du, dv: sum(du[k] * dv[k] for k in range(dim)))
        f = a.iasm(lambda v, x: F(*[x[k] for k in range(dim)]) * v)
        B = a.fasm(lambda u, v: u * v, normals=normals)
        g = a.fasm(lambda v, x: G(*[x[k] for k in range(dim)]) * v, normals=normals)
        u = spsolve((A + B).tocsc(), f + g)
        hs.append(mesh.param())
        errs.append(np.sqrt(a.L2error(u, U) ** 2 + a.H1error(u, dexact) ** 2))
    return np.polyfit(np.log10(hs), np.log10(errs), 1)[0]


@pytest.mark.gpu
def test_convergence_tetp2():               # test_module.py:82-157
    U = lambda x: 1 + x[0] - x[0] ** 2 * x[1] ** 2 + x[0] * x[1] * x[2] ** 3
    dexact = {0: lambda x: 1 - 2 * x[0] * x[1] ** 2 + x[1] * x[2] ** 3,
              1: lambda x: -2 * x[0] ** 2 * x[1] + x[0] * x[2] ** 3,
              2: lambda x: 3 * x[0] * x[1] * x[2] ** 2}
    F = lambda x, y, z: 2 * x ** 2 + 2 * y ** 2 - 6 * x * y * z
    G = lambda x, y, z: ((x == 1) * (3 - 3 * y ** 2 + 2 * y * z ** 3) + (x == 0) * (-y * z ** 3)
                         + (y == 1) * (1 + x - 3 * x ** 2 + 2 * x * z ** 3)
                         + (y == 0) * (1 + x - x * z ** 3)
                         + (z == 1) * (1 + x + 4 * x * y - x ** 2 * y ** 2)
                         + (z == 0) * (1 + x - x ** 2 * y ** 2))
    meshes = []
    for itr in range(1, 4):
        m = sp.MeshTet()
        m.refine(itr)
        meshes.append(m)
    rate = _robin_rate(meshes, sp.ElementTetP2, U, dexact, F, G)
    assert 1.95 <= rate <= 2.2


@pytest.mark.gpu
@pytest.mark.parametrize("ename,lo,hi", [("Q1", 0.95, 1.05), ("Q2", 1.95, 2.05)])
def test_convergence_q1_q2(ename, lo, hi):  # test_module.py:159-256
    U = lambda x: 1 + x[0] - x[0] ** 2 * x[1] ** 2 + np.exp(x[0])
    dexact = {0: lambda x: 1 - 2 * x[0] * x[1] ** 2 + np.exp(x[0]),
              1: lambda x: -2 * x[0] ** 2 * x[1]}
    F = lambda x, y: 2 * x ** 2 + 2 * y ** 2 - np.exp(x)
    G = lambda x, y: ((x == 1) * (3 - 3 * y ** 2 + 2 * np.exp(1)) + (x == 0) * 0
                      + (y == 1) * (1 + x - 3 * x ** 2 + np.exp(x)) + (y == 0) * (1 + x + np.exp(x)))
    meshes = []
    m = sp.MeshQuad()
    for _ in range(3):
        m.refine()
        meshes.append(sp.MeshQuad(m.p.copy(), m.t.copy()))
    rate = _robin_rate(meshes, getattr(sp, "Element" + ename), U, dexact, F, G, normals=False)
    assert lo <= rate <= hi
