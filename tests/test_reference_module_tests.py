# -*- coding: utf-8 -*-
"""The rest of the reference's unit tests that exercise the assembly path or its
input objects, restated against this package:

* spfem/test_asm.py:12-60     interior-facet (DG interior penalty) assembly
* spfem/test_module.py:397-473   TetP1Test (tetrahedral facets, error norms)
* spfem/test_module.py:534-642   ExamplePoisson (TetP1 / TetP2 convergence)
* spfem/test_module.py:644-711   ExampleElasticity (vector TetP1; the weak form
  that ``weakform.py`` generates there is written out by hand here)
* spfem/test_mapping.py:13-20    F / invF round trip
* spfem/test_mesh.py             faulty inputs, shape regularity, facet indexing

Every check runs twice: in the CPU suite with the generated kernels compiled
for the host (``on_host`` routes ``iasm`` / ``fasm`` and with them ``L2error`` /
``H1error`` through ``tests/hostemu.py``) and as a ``-m gpu`` test through the
public API."""
import copy

import numpy as np
import pytest
from scipy.sparse.linalg import spsolve

import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
import hostemu
from test_reference_known_answers import _Host


# ---------------------------------------------------------------- test_asm.py
def _interior_penalty(device):              # test_asm.py:12-60
    m = sp.MeshTri()
    m.refine(3)
    e = sp.element.ElementTriDG(sp.ElementTriP1())
    e1 = sp.ElementTriP1()
    wrap = (lambda a: a) if device else _Host
    a = wrap(AssemblerElement(m, e))
    a1 = wrap(AssemblerElement(m, e1))
    b1 = wrap(AssemblerElement(m, e1, e))
    A = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
    M = a.iasm(lambda u, v: u * v)
    b = a.iasm(lambda v, x: np.sin(2.0 * np.pi * x[0]) * v)
    B1 = a.fasm(lambda u1, u2, v1, v2, h: 1 / h * (u1 - u2) * (v1 - v2), interior=True)
    B2 = a.fasm(lambda u, v, h: 1 / h * u * v)
    B = B1 + B2
    # (the reference also assembles the consistency terms and multiplies them by 0)
    C1 = a.fasm(lambda du1, du2, v1, v2, n: 0.5 * (du1[0] * n[0] * v1 - du2[0] * n[0] * v2
                                                   + du1[1] * n[1] * v1 - du2[1] * n[1] * v2),
                interior=True)
    C2 = a.fasm(lambda du, v, n: (du[0] * n[0] + du[1] * n[1]) * v)
    C = C1 + C2
    D = C + C.T
    eta = 1e5
    x = sp.direct(A + eta * B - 0 * D, b)
    K = a1.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
    f = a1.iasm(lambda v, x: np.sin(2.0 * np.pi * x[0]) * v)
    y = sp.direct(K, f, I=m.interior_nodes())
    P = b1.iasm(lambda u, v: u * v)
    assert np.linalg.norm(x - sp.direct(M, P * y)) < 0.5e-4


def test_interior_penalty_host():
    _interior_penalty(False)


@pytest.mark.gpu
def test_interior_penalty_cuda():
    _interior_penalty(True)


# ------------------------------------------------------------- test_module.py
def _F(x, y, z):
    return 2 * x ** 2 + 2 * y ** 2 - 6 * x * y * z


def _G(x, y, z):
    return ((x == 1) * (3 - 3 * y ** 2 + 2 * y * z ** 3)
            + (x == 0) * (-y * z ** 3)
            + (y == 1) * (1 + x - 3 * x ** 2 + 2 * x * z ** 3)
            + (y == 0) * (1 + x - x * z ** 3)
            + (z == 1) * (1 + x + 4 * x * y - x ** 2 * y ** 2)
            + (z == 0) * (1 + x - x ** 2 * y ** 2))


def _exact(x):
    return 1 + x[0] - x[0] ** 2 * x[1] ** 2 + x[0] * x[1] * x[2] ** 3


_DEXACT = {0: lambda x: 1 - 2 * x[0] * x[1] ** 2 + x[1] * x[2] ** 3,
           1: lambda x: -2 * x[0] ** 2 * x[1] + x[0] * x[2] ** 3,
           2: lambda x: 3 * x[0] * x[1] * x[2] ** 2}


def _robin_tet(mesh, elem):
    """Solve the Robin problem of TetP1Test / ExamplePoisson; (h, L2, H1)."""
    a = AssemblerElement(mesh, elem)
    A = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1] + du[2] * dv[2])
    f = a.iasm(lambda v, x: _F(x[0], x[1], x[2]) * v)
    B = a.fasm(lambda u, v: u * v)
    g = a.fasm(lambda v, x: _G(x[0], x[1], x[2]) * v)
    u = spsolve((A + B).tocsc(), f + g)
    return mesh.param(), a.L2error(u, _exact), a.H1error(u, _DEXACT)


@pytest.fixture
def on_host(monkeypatch):
    """Route ``AssemblerElement.iasm / fasm`` (and with them ``L2error`` /
    ``H1error``) through the generated kernels compiled for the host."""
    monkeypatch.setattr(AssemblerElement, "iasm",
                        lambda self, form, **kw: hostemu.iasm(self, form, **kw))
    monkeypatch.setattr(AssemblerElement, "fasm",
                        lambda self, form, **kw: hostemu.fasm(self, form, **kw))


def _rate(rows):
    hs, l2, h1 = (np.array(c) for c in zip(*rows))
    return np.polyfit(np.log10(hs), np.log10(np.sqrt(l2 ** 2 + h1 ** 2)), 1)[0]


def _tetp1_refinements():                   # test_module.py:397-473
    rows = []
    for itr in range(2, 5):
        mesh = sp.MeshTet()
        mesh.refine(itr)
        rows.append(_robin_tet(mesh, sp.ElementTetP1()))
    assert _rate(rows) >= 1


def test_tetp1_refinements_host(on_host):
    _tetp1_refinements()


@pytest.mark.gpu
def test_tetp1_refinements():
    _tetp1_refinements()


def _example_poisson():                     # test_module.py:534-642
    mesh = sp.MeshTet()
    mesh.refine()
    r1, r2 = [], []
    for _ in range(3):
        r2.append(_robin_tet(mesh, sp.ElementTetP2()))
        mesh.refine()
        r1.append(_robin_tet(mesh, sp.ElementTetP1()))
    assert _rate(r1) >= 1
    assert _rate(r2) >= 2


def test_example_poisson_host(on_host):
    _example_poisson()


@pytest.mark.gpu
def test_example_poisson():
    _example_poisson()


def _elasticity_forms():
    """``dotp(Sigma(U), Eps(V))`` and ``dotp(div(Sigma(Ut)), V)`` of
    test_module.py:647-689 written out (du[i][j] = d u_i / d x_j)."""
    E, Nu = 20, 0.3
    lam = E * Nu / ((1 + Nu) * (1 - 2 * Nu))
    mu = E / (2 * (1 + Nu))

    def dudv(du, dv):
        out = lam * (du[0][0] + du[1][1] + du[2][2]) * (dv[0][0] + dv[1][1] + dv[2][2])
        for i in range(3):
            for j in range(3):
                out = out + 2 * mu * (0.5 * (du[i][j] + du[j][i])) * (0.5 * (dv[i][j] + dv[j][i]))
        return out

    pi = np.pi
    s, c = np.sin, np.cos

    def load(v, x):
        # Ut = (0, 0, w), w = 0.1 sin(pi x) sin(pi y) sin(pi z):
        # div Sigma(Ut) = ((lam+mu) w_xz, (lam+mu) w_yz, mu Lap(w) + (lam+mu) w_zz)
        X, Y, Z = pi * x[0], pi * x[1], pi * x[2]
        w = 0.1 * s(X) * s(Y) * s(Z)
        w_xz = 0.1 * pi ** 2 * c(X) * s(Y) * c(Z)
        w_yz = 0.1 * pi ** 2 * s(X) * c(Y) * c(Z)
        return ((lam + mu) * w_xz * v[0] + (lam + mu) * w_yz * v[1]
                + (mu * (-3 * pi ** 2 * w) + (lam + mu) * (-pi ** 2 * w)) * v[2])

    return dudv, load


def _elasticity(refine, device):
    dudv, load = _elasticity_forms()
    m = sp.MeshTet()
    m.refine(refine)
    e = sp.ElementH1Vec(sp.ElementTetP1())
    a = AssemblerElement(m, e)
    h = a if device else _Host(a)
    A = h.iasm(dudv)
    f = h.iasm(load)
    I = a.dofnum_u.getdofs(N=m.interior_nodes())
    u = sp.direct(A, f, I=I)
    return m, a, u


def test_example_elasticity_host():
    """Coarse version on the host emulation: symmetric stiffness matrix with
    the rigid-body modes in its kernel, displacement dominated by -sin sin sin
    in the z component."""
    m, a, u = _elasticity(2, False)
    uz = u[a.dofnum_u.n_dof[2, :]]
    exact = -0.1 * np.sin(np.pi * m.p[0]) * np.sin(np.pi * m.p[1]) * np.sin(np.pi * m.p[2])
    assert np.max(np.abs(uz - exact)) < 0.03
    assert np.max(np.abs(u[a.dofnum_u.n_dof[0, :]])) < 0.03
    dudv, _ = _elasticity_forms()
    A = hostemu.iasm(a, dudv)
    assert abs(A - A.T).max() < 1e-12
    for k in range(3):                       # translations
        t = np.zeros(A.shape[0])
        t[a.dofnum_u.n_dof[k, :]] = 1.0
        assert np.max(np.abs(A.dot(t))) < 1e-10
    rot = np.zeros(A.shape[0])               # rotation about z: (-y, x, 0)
    rot[a.dofnum_u.n_dof[0, :]] = -m.p[1]
    rot[a.dofnum_u.n_dof[1, :]] = m.p[0]
    assert np.max(np.abs(A.dot(rot))) < 1e-10


def _example_elasticity():                  # test_module.py:644-711
    m, a, u = _elasticity(4, True)
    c = AssemblerElement(m, sp.ElementTetP1())
    exact = lambda x: -0.1 * np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]) * np.sin(np.pi * x[2])
    assert c.L2error(u[a.dofnum_u.n_dof[2, :]], exact) <= 1e-3


def test_example_elasticity_full_host(on_host):
    _example_elasticity()


@pytest.mark.gpu
def test_example_elasticity():
    _example_elasticity()


# ------------------------------------------------------------ test_mapping.py
def test_mapping_F_invF_round_trip():       # test_mapping.py:13-20 (made strict)
    mesh = sp.MeshTri()
    mesh.refine()
    mapping = mesh.mapping()
    X = np.array([[0.1, 0.2, 0.3], [0.3, 0.2, 0.1]])
    y = mapping.F(X)
    Xb = mapping.invF(y)
    for k in range(2):
        assert np.allclose(Xb[k], np.tile(X[k], (mesh.t.shape[1], 1)), atol=1e-13)
    mt = sp.MeshTet()
    mt.refine()
    mp = mt.mapping()
    X3 = np.array([[0.1, 0.2], [0.3, 0.2], [0.2, 0.5]])
    Xb = mp.invF(mp.F(X3))
    for k in range(3):
        assert np.allclose(Xb[k], np.tile(X3[k], (mt.t.shape[1], 1)), atol=1e-13)


# --------------------------------------------------------------- test_mesh.py
def test_mesh_faulty_inputs():              # test_mesh.py:7-25
    with pytest.raises(Exception):           # point belonging to no element
        sp.MeshTri(p=np.array([[0, 0], [0, 1], [1, 0], [1, 1]]).T, t=np.array([[0, 1, 2]]).T)
    with pytest.raises(Exception):           # t not matching the mesh type
        sp.MeshTet(p=np.array([[0, 0], [0, 1], [1, 0], [1, 1]]).T, t=np.array([[0, 1, 2]]).T)
    with pytest.raises(Exception):
        sp.MeshLine(p=np.array([[0, 0], [0, 1], [1, 0], [1, 1]]).T, t=np.array([[0, 1, 2]]).T)
    with pytest.raises(Exception):           # transposed inputs
        sp.MeshTri(p=np.array([[0, 0], [0, 1], [1, 0], [1, 1]]),
                   t=np.array([[0, 1, 2], [1, 2, 3]]))


def test_mesh_tet_refine_shape_regularity():    # test_mesh.py:28-35
    m = sp.MeshTet()
    for _ in range(4):                       # (5 levels in the reference; 4 keep the CPU suite short)
        m.refine()
        assert m.shapereg() < 2.5


def test_mesh_tri_indexing():               # test_mesh.py:44-77
    base = sp.MeshTri()
    base.refine()
    mesh = copy.deepcopy(base)
    mesh.refine(4)
    assert np.max(mesh.t) == mesh.p.shape[1] - 1
    mesh = copy.deepcopy(base)
    mesh.refine(2)
    nf = mesh.facets.shape[1]
    hit = np.equal(mesh.t2f[:, mesh.f2t[0, :]], np.tile(np.arange(nf), (3, 1)))
    assert np.sum(hit) == nf
    # walking t2f / f2t from one triangle reaches the whole mesh
    cur = np.array([4])
    for _ in range(15):
        add = np.array([], dtype=np.int64)
        for j in cur:
            if j != -1:
                add = np.append(add, np.unique(mesh.f2t[:, mesh.t2f[:, j]].flatten()))
        cur = np.unique(np.append(cur, add))
    assert cur.shape[0] - 1 == mesh.t.shape[1]
