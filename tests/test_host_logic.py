# -*- coding: utf-8 -*-
"""Host logic without a GPU: API surface / error behaviour of the reference,
tracer rules, NVRTC compilation for sm_100a, and the C-ABI exports."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import spfem_b200 as sp
from spfem_b200 import _lib
from spfem_b200.assembly import AssemblerElement, Dofnum
from spfem_b200.tracer import TraceError
from spfem_b200.engine import scipy_index_dtype
import cases
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def test_c_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "spfem_b200.h")).read()
    declared = set(re.findall(r"\b(spfem_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r"\bT (spfem_[a-z_0-9]+)", out))
    assert declared <= exported, declared - exported
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.spfem_abi_version() == 4
    assert lib.spfem_device_count() >= 0


def test_nvrtc_compiles_generated_kernels_for_sm100a(lib):
    for name in ("trip1_lap", "vecp2_elasticity", "q2_lap", "f_vecp2_nitsche",
                 "f_trip1_sipg", "trip1_interp_mass"):
        case = cases.BY_NAME[name]
        asm = util.assembler_for(case)
        kw = cases.case_inputs(case, asm.mesh, asm.dofnum_u.N)
        kw.pop("tind", None)
        kw.pop("find", None)
        plan = (asm.plan_iasm(case.form, **kw) if case.kind == "iasm"
                else asm.plan_fasm(case.form, **kw))
        image = _lib.nvrtc_compile(plan.source)
        assert image[:4] == b"\x7fELF" and len(image) > 1000


def test_nvrtc_error_is_reported(lib):
    with pytest.raises(_lib.SpfemError):
        _lib.nvrtc_compile("this is not CUDA")


def test_constructor_errors():
    m = sp.MeshTri()
    with pytest.raises(Exception, match="First parameter must be an instance"):
        AssemblerElement("mesh", sp.ElementTriP1())
    with pytest.raises(Exception, match="Second parameter must be an instance"):
        AssemblerElement(m, "elem")


def test_interp_must_be_dict_and_unknown_argument():
    m = sp.MeshTri()
    m.refine(1)
    a = AssemblerElement(m, sp.ElementTriP1())
    with pytest.raises(Exception, match="must be in a dictionary"):
        a.plan_iasm(lambda v, w: w[0] * v, interp=np.zeros(3))
    with pytest.raises(IndexError):
        a.plan_iasm(lambda v, foo: v)
    with pytest.raises(Exception, match="No interior support"):
        a.plan_fasm(lambda v1, v2: v1, interior=True)


def test_nonlinear_forms_are_rejected():
    m = sp.MeshTri()
    a = AssemblerElement(m, sp.ElementTriP1())
    with pytest.raises(TraceError):
        a.plan_iasm(lambda u, v: u * u * v)
    with pytest.raises(TraceError):
        a.plan_iasm(lambda u, v: np.sin(u) * v)
    with pytest.raises(TraceError):
        a.plan_iasm(lambda v, x: v if x[0] > 0 else 0 * v)


def test_dofnum_matches_oracle():
    import asm_oracle
    for mesh, ref, el, nc in (("MeshTri", 3, "TriP2", 1), ("MeshTri", 2, "TriP2", 2),
                              ("MeshTet", 1, "TetP2", 1), ("MeshQuad", 2, "Q2", 1),
                              ("MeshTet", 1, "TetP1", 3)):
        m = getattr(sp, mesh)()
        m.refine(ref)
        d = Dofnum(m, util.element(el, nc))
        t_dof, N = asm_oracle.dofnum(m, el, nc)
        assert d.N == N and np.array_equal(d.t_dof, t_dof)


def test_scipy_index_width_rule():
    assert scipy_index_dtype(10, 5, 5) == np.int32
    assert scipy_index_dtype(2 ** 31 - 1, 5, 5) == np.int32
    assert scipy_index_dtype(2 ** 31, 5, 5) == np.int64
    assert scipy_index_dtype(10, 2 ** 31, 5) == np.int64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = sp.MeshTri()
    a = AssemblerElement(m, sp.ElementTriP1())
    with pytest.raises(_lib.SpfemError, match="no CPU fallback"):
        a.iasm(cases.mass)


def test_tensor_mode_precontracts_quadrature():
    """Element-constant coefficients fold the quadrature sum into constants:
    the P1 Laplace kernel has no per-point code at all, and symmetric /
    repeated local entries are computed once (6 distinct values for the P1
    stiffness matrix; the P1 mass matrix tables differ in the last bit, so only
    exactly equal entries are shared)."""
    m = sp.MeshTri()
    a = AssemblerElement(m, sp.ElementTriP1())
    plan = a.plan_iasm(cases.lap2)
    assert "x0_" not in plan.source
    assert plan.spec.n_entries == 9 and plan.spec.n_unique == 6
    assert plan.spec.alias[1] == plan.spec.alias[3]          # A[0,1] == A[1,0]
    assert a.plan_iasm(cases.mass).spec.n_unique <= 6


@pytest.mark.parametrize("name", ["trip1_lap_r6", "tetp2_lap", "vecp2_elasticity", "stokes_b",
                                  "trip1_load", "trip1_interp_mass", "trip1_interp_load"])
def test_chunked_assembly_matches_single_call(name, monkeypatch):
    """Meshes with more local entries than one pattern call takes are assembled
    chunk by chunk (AssemblerElement._iasm_chunked); the union of the owned rows
    is the reference matrix, pattern included (generated kernels on the host)."""
    import hostemu
    case = cases.BY_NAME[name] if name in cases.BY_NAME else None
    if case is None:
        pytest.skip("no such case")
    asm = util.assembler_for(case)
    kw = cases.case_inputs(case, asm.mesh, asm.dofnum_u.N)
    interp = kw.get("interp")                    # interpolated fields follow the chunks' DOFs
    plan = asm.plan_iasm(case.form, interp=interp)
    n_ent = plan.spec.n_entries * asm.mesh.t.shape[1]
    monkeypatch.setenv("SPFEM_MAX_ENTRIES", str(n_ent // 3))
    chunks = asm._chunks_needed(plan, None)
    assert chunks >= 4
    out = asm._iasm_chunked(case.form, plan, None, interp, chunks,
                            assemble=lambda loc, form, **kw: hostemu.iasm(loc, form, **kw))
    util.check(out, util.golden(case))
    monkeypatch.setenv("SPFEM_MAX_ENTRIES", str(2 ** 31 - 1))
    assert asm._chunks_needed(plan, None) == 1


def test_spfem_alias_package():
    """``compat/spfem``: reference-style imports resolve to spfem_b200
    (spfem/__init__.py:1-6, spfem/asm.py:7)."""
    code = ("import spfem, spfem.mesh, spfem.asm, spfem.element\n"
            "from spfem.asm import AssemblerElement\n"
            "from spfem.mesh import MeshTri\n"
            "import spfem_b200.assembly as a\n"
            "assert AssemblerElement is a.AssemblerElement and spfem.asm is a\n"
            "m = MeshTri(); m.refine(2)\n"
            "assert spfem.AssemblerElement(m, spfem.ElementTriP1()).dofnum_u.N == 25\n"
            "print('alias ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "compat") + os.pathsep + ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "alias ok" in out.stdout, out.stderr[-2000:]


def test_captured_scalars_are_runtime_parameters():
    """A Python scalar captured by the form (time step, penalty ...) must not be
    baked into the kernel source: two values -> the same source (one NVRTC
    compile), different ``params``; structural constants stay literals."""
    import hostemu
    m = sp.MeshTri()
    m.refine(2)
    a = AssemblerElement(m, sp.ElementTriP1())

    def make(dt, kappa):
        return lambda u, v, du, dv: u * v + dt * kappa * (du[0] * dv[0] + du[1] * dv[1]) + 0.5 * u * v

    p1 = a.plan_iasm(make(0.013, 3.7))
    p2 = a.plan_iasm(make(0.029, 1.1))
    assert p1.source == p2.source
    assert p1.spec.params != p2.spec.params and len(p1.spec.params) == 1
    assert "0x1.0000000000000p-1" in p1.source or "0.5" in p1.source or "p-1" in p1.source
    M = hostemu.iasm(a, lambda u, v: u * v)
    K = hostemu.iasm(a, lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
    for dt, kappa in ((0.013, 3.7), (0.029, 1.1)):
        A = hostemu.iasm(a, make(dt, kappa))
        ref = 1.5 * M + dt * kappa * K
        assert abs(A - ref).max() < 1e-13
    # facet forms as well
    f1 = a.plan_fasm(lambda u, v, h: 17.3 / h * u * v)
    f2 = a.plan_fasm(lambda u, v, h: 41.9 / h * u * v)
    assert f1.source == f2.source and f1.spec.params == [17.3] and f2.spec.params == [41.9]


def test_foreign_mapping_is_rejected():
    """A mapping that is not the mesh's own would be ignored by the device
    kernels (spfem/assembly.py:829-832 uses whatever mapping it is given)."""
    m = sp.MeshTri()
    m.refine(1)
    AssemblerElement(m, sp.ElementTriP1(), mapping=m.mapping())          # the mesh's own map: fine
    other = sp.MeshTri()
    other.refine(1)
    other.p[0] *= 2.0
    with pytest.raises(NotImplementedError):
        AssemblerElement(m, sp.ElementTriP1(), mapping=other.mapping())
    q = sp.MeshQuad()
    AssemblerElement(q, sp.ElementQ1(), mapping=q.mapping())
