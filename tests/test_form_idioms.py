# -*- coding: utf-8 -*-
"""Weak-form idioms of the reference tree that go beyond
plain arithmetic on the arguments: they must trace, and the traced kernel must
give what the reference's NumPy evaluation of the same Python function gives
(the oracle calls the form on arrays, exactly like spfem/assembly.py:961-984).

* ``copy.deepcopy(dw)`` + item assignment (symmetric gradient of the Stokes
  notebook, learning/Example 1 - Stokes equations.ipynb cells 9-10),
* ``np.array([[dU[0][0], ...]])`` of the arguments and ``T[0, 0]`` indexing
  (tmp/elastic_contact.py:49-72),
* ``eval``'d lambdas with the fixed signature ``u,v,du,dv,x`` as produced by
  ``TensorFunction.handlify`` (spfem/weakform.py:235-288, used by
  spfem/test_module.py:668-698)."""
import copy

import numpy as np
import pytest

import cases
import hostemu
import util
import asm_oracle


# -- learning/Example 1 - Stokes equations.ipynb, cell 9 ---------------------
def stokes_bilinear_a(du, dv):
    def inner_product(a, b):
        return a[0][0]*b[0][0] +\
               a[0][1]*b[0][1] +\
               a[1][0]*b[1][0] +\
               a[1][1]*b[1][1]

    def eps(dw):  # symmetric part of the velocity gradient
        dW = copy.deepcopy(dw)
        dW[0][1] = .5*(dw[0][1] + dw[1][0])
        dW[1][0] = dW[0][1]
        return dW
    return inner_product(eps(du), eps(dv))


def stokes_bilinear_b(du, v):
    return (du[0][0]+du[1][1])*v


# -- tmp/elastic_contact.py:40-72 -------------------------------------------
E1, nu1 = 1.0, 0.3
mu1 = E1/(2.0*(1.0+nu1))
lam1 = nu1*E1/((1.0+nu1)*(1.0-2.0*nu1))


def Eps(dU):
    return np.array([[dU[0][0], .5*(dU[0][1]+dU[1][0])],
                     [.5*(dU[1][0]+dU[0][1]), dU[1][1]]])


def ddot(T1, T2):
    return T1[0, 0]*T2[0, 0] +\
           T1[0, 1]*T2[0, 1] +\
           T1[1, 0]*T2[1, 0] +\
           T1[1, 1]*T2[1, 1]


def C1(T):
    trT = T[0, 0]+T[1, 1]
    return np.array([[2.*mu1*T[0, 0]+lam1*trT, 2.*mu1*T[0, 1]],
                     [2.*mu1*T[1, 0], 2.*mu1*T[1, 1]+lam1*trT]])


def dudv1(du, dv):
    return ddot(C1(Eps(du)), Eps(dv))


# -- what TensorFunction.handlify returns (spfem/weakform.py:279-288): a string
#    over u, v, du[i][j], dv[i][j], x[k], np.sin / np.cos / np.sqrt / np.pi,
#    eval'd into a lambda with the FIXED argument list u,v,du,dv,x ------------
HANDLIFY_BILINEAR = eval(
    "lambda u,v,du,dv,x: "
    "7.69230769230769*(du[0][1] + du[1][0])*(dv[0][1] + dv[1][0])/2"
    " + 11.5384615384615*(du[0][0] + du[1][1])*(dv[0][0] + dv[1][1])"
    " + 15.3846153846154*du[0][0]*dv[0][0] + 15.3846153846154*du[1][1]*dv[1][1]")
HANDLIFY_LOAD = eval(
    "lambda v,dv,x: "
    "0.1*np.pi**2*np.sin(np.pi*x[0])*np.cos(np.pi*x[1])*v[0]"
    " - np.sqrt(x[0]**2 + x[1]**2 + 1)*v[1]")


IDIOMS = {
    # name: (form, mesh, refine, trial element, components u, test element, components v)
    "stokes_notebook_A": (stokes_bilinear_a, "MeshTri", 2, "TriP2", 2, None, 2),
    "stokes_notebook_B": (stokes_bilinear_b, "MeshTri", 2, "TriP2", 2, "TriP1", 1),
    "elastic_contact_nparray": (dudv1, "MeshTri", 3, "TriP1", 2, None, 2),
    "handlify_bilinear": (HANDLIFY_BILINEAR, "MeshTri", 2, "TriP2", 2, None, 2),
    "handlify_load": (HANDLIFY_LOAD, "MeshTri", 3, "TriP1", 2, None, 2),
}


def _setup(name):
    form, mesh, refine, eu, ncu, ev, ncv = IDIOMS[name]
    case = cases.Case(name, mesh, refine, eu, form, ncu=ncu, ev=ev, ncv=ncv)
    asm = util.assembler_for(case)
    ref = asm_oracle.iasm(asm.mesh, eu, form, elem_v=ev, ncomp_u=ncu, ncomp_v=ncv)
    return asm, form, ref


@pytest.mark.parametrize("name", sorted(IDIOMS))
def test_idiom_traces_and_matches_oracle_host(name):
    asm, form, ref = _setup(name)
    util.check(hostemu.iasm(asm, form), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(IDIOMS))
def test_idiom_matches_oracle_cuda(name):
    asm, form, ref = _setup(name)
    util.check(asm.iasm(form), ref)


def test_deepcopy_form_equals_written_out_form():
    """The notebook's deepcopy form is the sym-grad form written out."""
    asm, form, ref = _setup("stokes_notebook_A")

    def written_out(du, dv):
        e01u = .5*(du[0][1] + du[1][0])
        e01v = .5*(dv[0][1] + dv[1][0])
        return du[0][0]*dv[0][0] + 2*e01u*e01v + du[1][1]*dv[1][1]
    B = asm_oracle.iasm(asm.mesh, "TriP2", written_out, ncomp_u=2, ncomp_v=2)
    assert np.max(np.abs((ref - B).data)) <= 1e-13 * np.abs(ref.data).max()
