# -*- coding: utf-8 -*-
"""``fnorm`` (facet-wise norms of interpolated solutions, spfem/assembly.py:1185-1271):
generated facet kernel on the host and on the GPU vs the reference's own output
(tests/golden/fnorm_*.npz, written by tests/golden/make_golden.py)."""
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
    interp = {0: cases._interp_vec(a.dofnum_u.N, 3)}
    gold = np.load(os.path.join(util.GOLDEN, "fnorm_%s_%s.npz" % (mname, ename)))
    return a, interp, gold


def _check(val, find, gold, which):
    assert np.array_equal(find, gold[which + "_find"])
    g = gold[which]
    assert val.shape == g.shape
    assert np.max(np.abs(val - g)) <= 1e-12 * np.max(np.abs(g))


@pytest.mark.parametrize("mname,ename,ref", cases.FNORM_CASES)
def test_fnorm_generated_kernel(mname, ename, ref):
    a, interp, gold = _setup(mname, ename, ref)
    val, find = hostemu.fnorm(a, cases.fnorm_boundary_form, interp)
    _check(val, find, gold, "boundary")
    val, find = hostemu.fnorm(a, cases.fnorm_interior_form, interp, interior=True)
    _check(val, find, gold, "interior")


@pytest.mark.gpu
@pytest.mark.parametrize("mname,ename,ref", cases.FNORM_CASES)
def test_fnorm_cuda(mname, ename, ref):
    a, interp, gold = _setup(mname, ename, ref)
    val, find = a.fnorm(cases.fnorm_boundary_form, interp)
    _check(val, find, gold, "boundary")
    val, find = a.fnorm(cases.fnorm_interior_form, interp, interior=True)
    _check(val, find, gold, "interior")


def test_fnorm_errors():
    a, interp, _ = _setup("MeshTri", "TriP1", 2)
    with pytest.raises(Exception, match="must be in a dictionary"):
        a.fnorm(cases.fnorm_boundary_form, interp[0])
    with pytest.raises(Exception, match="no interior=True"):
        a.fnorm(cases.fnorm_interior_form, interp)
