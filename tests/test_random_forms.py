# -*- coding: utf-8 -*-
"""Differential test of the tracer + code generator: randomly composed weak
forms (sums of products of ``u, v, du, dv`` with coefficients built from
``x``, ``h``, an interpolated field ``w`` and NumPy functions) assembled by the
generated kernel -- compiled for the host here, on the GPU with ``-m gpu`` --
against the NumPy oracle, which evaluates the very same Python callable on
arrays like the reference does (spfem/assembly.py:971-983)."""
import numpy as np
import pytest

import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
import asm_oracle
import cases
import hostemu

MESHES = {
    "TriP1": lambda: _refined(sp.MeshTri, 2), "TriP2": lambda: _refined(sp.MeshTri, 1),
    "TetP1": lambda: _refined(sp.MeshTet, 1), "Q1": lambda: _refined(sp.MeshQuad, 2),
}


def _refined(cls, n):
    m = cls()
    m.refine(n)
    return m


def _coefficient(rng, dim, with_w):
    """A random scalar coefficient as ``(function of (x, h, w), text)``."""
    kinds = ["const", "x", "sinx", "expx", "h", "mask", "xx"] + (["w"] if with_w else [])
    k = kinds[rng.integers(len(kinds))]
    a = float(np.round(rng.uniform(0.5, 2.0), 3))
    i = int(rng.integers(dim))
    j = int(rng.integers(dim))
    if k == "const":
        return (lambda x, h, w: a), "%g" % a
    if k == "x":
        return (lambda x, h, w: a + x[i]), "(%g+x%d)" % (a, i)
    if k == "sinx":
        return (lambda x, h, w: np.sin(a * x[i]) + 2.0), "(sin(%g x%d)+2)" % (a, i)
    if k == "expx":
        return (lambda x, h, w: np.exp(-a * x[i] * x[j])), "exp(-%g x%d x%d)" % (a, i, j)
    if k == "h":
        return (lambda x, h, w: a / h + h ** 2), "(%g/h+h^2)" % a
    if k == "mask":
        return (lambda x, h, w: 1.0 + a * (x[i] >= 0.5) * (x[j] <= 0.75)), "mask%d%d" % (i, j)
    if k == "xx":
        return (lambda x, h, w: x[i] ** 2 + np.sqrt(1.0 + x[j])), "(x%d^2+sqrt(1+x%d))" % (i, j)
    return (lambda x, h, w: 1.0 + w[0] ** 2), "(1+w^2)"


def _random_form(rng, dim, bilinear, with_w):
    """Sum of 1-3 terms ``coef * (trial factor) * (test factor)``."""
    terms, text = [], []
    for _ in range(int(rng.integers(1, 4))):
        c, ct = _coefficient(rng, dim, with_w)
        tu = int(rng.integers(dim + 1))            # 0: value, 1 + k: derivative k
        tv = int(rng.integers(dim + 1))
        terms.append((c, tu, tv))
        text.append("%s*%s*%s" % (ct, "u" if tu == 0 else "du%d" % (tu - 1),
                                  "v" if tv == 0 else "dv%d" % (tv - 1)))

    def pick(val, grad, t):
        return val if t == 0 else grad[t - 1]
    if bilinear and with_w:
        def form(u, v, du, dv, x, w, h):
            return sum(c(x, h, w) * pick(u, du, tu) * pick(v, dv, tv) for c, tu, tv in terms)
    elif bilinear:
        def form(u, v, du, dv, x, h):
            return sum(c(x, h, None) * pick(u, du, tu) * pick(v, dv, tv) for c, tu, tv in terms)
    elif with_w:
        def form(v, dv, x, w, h):
            return sum(c(x, h, w) * pick(v, dv, tv) for c, tu, tv in terms)
    else:
        def form(v, dv, x, h):
            return sum(c(x, h, None) * pick(v, dv, tv) for c, tu, tv in terms)
    return form, " + ".join(text)


def _cases():
    out = []
    rng = np.random.default_rng(20240)
    for ename in ("TriP1", "TriP2", "TetP1", "Q1"):
        for k in range(4):
            out.append((ename, k, bool(k % 2 == 0), bool(k >= 2), int(rng.integers(1 << 30))))
    return out


def _run(ename, bilinear, with_w, seed, device):
    mesh = MESHES[ename]()
    dim = mesh.p.shape[0]
    rng = np.random.default_rng(seed)
    form, text = _random_form(rng, dim, bilinear, with_w)
    a = AssemblerElement(mesh, getattr(sp, "Element" + ename)())
    interp = {0: np.cos(3.0 * np.arange(a.dofnum_u.N))} if with_w else None
    ref = asm_oracle.iasm(mesh, ename, form, interp=interp)
    got = (a.iasm(form, interp=interp) if device else hostemu.iasm(a, form, interp=interp))
    if bilinear:
        cases.compare_csr(got, ref)
    else:
        cases.compare_vec(got, ref)


@pytest.mark.parametrize("ename,k,bilinear,with_w,seed", _cases())
def test_random_form_host(ename, k, bilinear, with_w, seed):
    _run(ename, bilinear, with_w, seed, False)


@pytest.mark.gpu
@pytest.mark.parametrize("ename,k,bilinear,with_w,seed", _cases())
def test_random_form_cuda(ename, k, bilinear, with_w, seed):
    _run(ename, bilinear, with_w, seed, True)


# ---------------------------------------------------------------------------
# boundary facet forms: the same idea with normals and the facet size
# ---------------------------------------------------------------------------
def _random_facet_form(rng, dim, bilinear):
    terms = []
    for _ in range(int(rng.integers(1, 4))):
        c, _t = _coefficient(rng, dim, False)
        tu = int(rng.integers(dim + 1))
        tv = int(rng.integers(dim + 1))
        kn = int(rng.integers(dim))                # multiply by a normal component
        terms.append((c, tu, tv, kn))

    def pick(val, grad, t):
        return val if t == 0 else grad[t - 1]
    if bilinear:
        def form(u, v, du, dv, x, h, n):
            return sum(c(x, h, None) * n[kn] * pick(u, du, tu) * pick(v, dv, tv)
                       for c, tu, tv, kn in terms)
    else:
        def form(v, dv, x, h, n):
            return sum(c(x, h, None) * n[kn] * pick(v, dv, tv) for c, tu, tv, kn in terms)
    return form


def _facet_cases():
    rng = np.random.default_rng(777)
    return [(en, bool(k % 2 == 0), int(rng.integers(1 << 30)))
            for en in ("TriP1", "TriP2", "TetP1") for k in range(2)]


def _run_facet(ename, bilinear, seed, device):
    mesh = MESHES[ename]()
    form = _random_facet_form(np.random.default_rng(seed), mesh.p.shape[0], bilinear)
    a = AssemblerElement(mesh, getattr(sp, "Element" + ename)())
    ref = asm_oracle.fasm(mesh, ename, form, normals=True)
    got = a.fasm(form, normals=True) if device else hostemu.fasm(a, form, normals=True)
    (cases.compare_csr if bilinear else cases.compare_vec)(got, ref)


@pytest.mark.parametrize("ename,bilinear,seed", _facet_cases())
def test_random_facet_form_host(ename, bilinear, seed):
    _run_facet(ename, bilinear, seed, False)


@pytest.mark.gpu
@pytest.mark.parametrize("ename,bilinear,seed", _facet_cases())
def test_random_facet_form_cuda(ename, bilinear, seed):
    _run_facet(ename, bilinear, seed, True)


# ---------------------------------------------------------------------------
# vector elements (ElementH1Vec): random bilinear forms in the components
# ---------------------------------------------------------------------------
def _random_vector_form(rng, dim):
    terms = []
    for _ in range(int(rng.integers(2, 5))):
        c, _t = _coefficient(rng, dim, False)
        cu, cv = int(rng.integers(dim)), int(rng.integers(dim))
        tu, tv = int(rng.integers(dim + 1)), int(rng.integers(dim + 1))
        terms.append((c, cu, cv, tu, tv))

    def pick(val, grad, comp, t):
        return val[comp] if t == 0 else grad[comp][t - 1]

    def form(u, v, du, dv, x, h):
        return sum(c(x, h, None) * pick(u, du, cu, tu) * pick(v, dv, cv, tv)
                   for c, cu, cv, tu, tv in terms)
    return form


@pytest.mark.parametrize("ename,seed", [("TriP1", 11), ("TriP2", 12), ("TetP1", 13), ("TriP1", 14)])
def test_random_vector_form_host(ename, seed):
    mesh = MESHES[ename]()
    dim = mesh.p.shape[0]
    form = _random_vector_form(np.random.default_rng(seed), dim)
    a = AssemblerElement(mesh, sp.ElementH1Vec(getattr(sp, "Element" + ename)()))
    ref = asm_oracle.iasm(mesh, ename, form, ncomp_u=dim, ncomp_v=dim)
    cases.compare_csr(hostemu.iasm(a, form), ref)


# ---------------------------------------------------------------------------
# interior facets: two-sided forms (u1, u2, v1, v2, du1, ...) as used for DG / jumps
# ---------------------------------------------------------------------------
def _random_interior_form(rng, dim):
    terms = []
    for _ in range(int(rng.integers(2, 5))):
        c, _t = _coefficient(rng, dim, False)
        su, sv = int(rng.integers(2)), int(rng.integers(2))     # side of trial / test
        tu, tv = int(rng.integers(dim + 1)), int(rng.integers(dim + 1))
        kn = int(rng.integers(dim))
        terms.append((c, su, sv, tu, tv, kn))

    def pick(val, grad, t):
        return val if t == 0 else grad[t - 1]

    def form(u1, u2, v1, v2, du1, du2, dv1, dv2, x, h, n):
        U, V, dU, dV = (u1, u2), (v1, v2), (du1, du2), (dv1, dv2)
        return sum(c(x, h, None) * n[kn] * pick(U[su], dU[su], tu) * pick(V[sv], dV[sv], tv)
                   for c, su, sv, tu, tv, kn in terms)
    return form


@pytest.mark.parametrize("ename,seed", [("TriP1", 21), ("TriP2", 22), ("TetP1", 23)])
def test_random_interior_facet_form_host(ename, seed):
    mesh = MESHES[ename]()
    form = _random_interior_form(np.random.default_rng(seed), mesh.p.shape[0])
    a = AssemblerElement(mesh, getattr(sp, "Element" + ename)())
    ref = asm_oracle.fasm(mesh, ename, form, interior=True, normals=True)
    cases.compare_csr(hostemu.fasm(a, form, interior=True, normals=True), ref)
