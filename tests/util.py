# -*- coding: utf-8 -*-
"""Builders shared by the test modules."""
import os

import numpy as np
from scipy.sparse import csr_matrix

import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_MESH_CACHE = {}


def mesh_for(case):
    key = (case.mesh, case.refine)
    if key not in _MESH_CACHE:
        m = getattr(sp, case.mesh)()
        m.refine(case.refine)
        _MESH_CACHE[key] = m
    return _MESH_CACHE[key]


def element(name, ncomp):
    if name == "TriDG1":
        e = sp.element.ElementTriDG(sp.element.ElementTriP1())
    elif name.startswith("TriPp"):
        e = sp.element.ElementTriPp(int(name[5:]))
    else:
        e = getattr(sp.element, "Element" + name)()
    return sp.ElementH1Vec(e) if ncomp > 1 else e


def assembler_for(case):
    m = mesh_for(case)
    eu = element(case.eu, case.ncu)
    if case.ev is None and case.ncv == case.ncu:
        return AssemblerElement(m, eu)
    return AssemblerElement(m, eu, element(case.ev or case.eu, case.ncv))


def golden(case):
    z = np.load(os.path.join(GOLDEN, case.name + ".npz"))
    if str(z["kind"]) == "csr":
        return csr_matrix((z["data"], z["indices"], z["indptr"]),
                          shape=tuple(z["shape"]))
    return z["data"]


def oracle_run(case):
    import asm_oracle
    m = mesh_for(case)
    _, nu = asm_oracle.dofnum(m, case.eu, case.ncu)
    kw = cases.case_inputs(case, m, nu)
    fn = asm_oracle.iasm if case.kind == "iasm" else asm_oracle.fasm
    return fn(m, case.eu, case.form, elem_v=case.ev, ncomp_u=case.ncu,
              ncomp_v=case.ncv, **kw)


def check(out, ref):
    if hasattr(ref, "indptr"):
        return cases.compare_csr(out, ref)
    return cases.compare_vec(out, ref)
