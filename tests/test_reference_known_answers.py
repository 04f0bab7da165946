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
