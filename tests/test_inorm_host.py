# -*- coding: utf-8 -*-
"""``inorm`` (element-wise norms of interpolated solutions and their gradients):
generated kernel on the host vs the reference's own output."""
import os

import numpy as np
import pytest

import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
import cases
import util
import hostemu


def _setup(mname, ename, ref):
    m = getattr(sp, mname)()
    m.refine(ref)
    a = AssemblerElement(m, util.element(ename, 1))
    N = a.dofnum_u.N
    interp = {0: cases._interp_vec(N, 3)}
    form = cases.inorm_form
    if m.p.shape[0] == 3:
        form = lambda u, du, x, h: cases.inorm_form(u, du, x, h) + du[0][2]
    gold = np.load(os.path.join(util.GOLDEN, "inorm_%s_%s.npz" % (mname, ename)))["data"]
    return a, form, interp, gold


@pytest.mark.parametrize("mname,ename,ref", cases.INORM_CASES)
def test_inorm_generated_kernel(mname, ename, ref):
    a, form, interp, gold = _setup(mname, ename, ref)
    val = hostemu.inorm(a, form, interp)
    assert val.shape == gold.shape
    assert np.max(np.abs(val - gold)) <= 1e-12 * np.max(np.abs(gold))


@pytest.mark.gpu
@pytest.mark.parametrize("mname,ename,ref", cases.INORM_CASES)
def test_inorm_cuda(mname, ename, ref):
    a, form, interp, gold = _setup(mname, ename, ref)
    val = a.inorm(form, interp)
    assert np.max(np.abs(val - gold)) <= 1e-12 * np.max(np.abs(gold))
    with pytest.raises(Exception, match="must be in a dictionary"):
        a.inorm(form, interp[0])


def test_inorm_two_fields_against_direct_quadrature():
    """Several interpolated fields (not comparable with the reference, which
    aliases them): direct NumPy quadrature of the same integrand."""
    from spfem_b200.quadrature import get_quadrature
    m = sp.MeshTri()
    m.refine(3)
    a = AssemblerElement(m, sp.ElementTriP1())
    N = a.dofnum_u.N
    interp = {"a": cases._interp_vec(N, 3), "b": cases._interp_vec(N, 4)}
    val = hostemu.inorm(a, lambda u, du: u["a"] * du["b"][1] - u["b"], interp)
    X, W = get_quadrature("tri", 2)
    phi = np.array([1 - X[0] - X[1], X[0], X[1]])
    dphi = np.array([[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]])          # reference gradients
    mp = m.mapping()
    td = a.dofnum_u.t_dof
    wa = np.einsum("jt,jq->tq", interp["a"][td], phi)
    wb = np.einsum("jt,jq->tq", interp["b"][td], phi)
    gb = [np.einsum("jt,j->t", interp["b"][td], dphi[l]) for l in range(2)]
    dwb1 = mp.invA[0][1] * gb[0] + mp.invA[1][1] * gb[1]
    expect = ((wa * dwb1[:, None] - wb) ** 2 * np.abs(mp.detA)[:, None]) @ W
    assert np.max(np.abs(val - expect)) <= 1e-12 * np.max(np.abs(expect))
