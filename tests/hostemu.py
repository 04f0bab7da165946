# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE: run a *generated* kernel on the CPU.

The code generator emits the element body as a ``SPFEM_DEVICE`` function; with
``-DSPFEM_HOST`` the identical text compiles with g++ (``-ffp-contract=off`` so
the ``__dmul_rn`` style intrinsics keep their no-FMA meaning).  This lets the
CPU test-suite check tracer + code generator numerics against the oracle
without a GPU.  It is NOT a product path: nothing under ``spfem_b200/``
imports this module, and the duplicate summation here is SciPy's."""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np
from scipy.sparse import coo_matrix

from spfem_b200._lib import SpfemArgs

_CACHE = {}
_TMP = os.path.join(tempfile.gettempdir(), "spfem_hostemu")


def compile_host(plan):
    key = hashlib.sha256(plan.source.encode()).hexdigest()[:20]
    if key in _CACHE:
        return _CACHE[key]
    os.makedirs(_TMP, exist_ok=True)
    src = os.path.join(_TMP, key + ".cpp")
    so = os.path.join(_TMP, key + ".so")
    if not os.path.exists(so):
        # private temporaries + atomic rename: several test processes (gloo ranks,
        # xdist workers) may compile the same source at the same time
        tag = "%s.%d" % (key, os.getpid())
        src = os.path.join(_TMP, tag + ".cpp")
        tmp_so = os.path.join(_TMP, tag + ".so.tmp")
        with open(src, "w") as fh:
            fh.write(plan.source)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-DSPFEM_HOST",
                               "-shared", "-fPIC", "-o", tmp_so, src])
        os.replace(tmp_so, so)
        os.remove(src)
    lib = C.CDLL(so)
    fn = getattr(lib, plan.name + "_host")
    fn.argtypes = [C.POINTER(SpfemArgs)]
    fn.restype = None
    _CACHE[key] = fn
    return fn


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def run_interior(asm, plan, tind=None, interp=None):
    """Local entries ``out[entry, k]`` of the interior kernel."""
    mesh = asm.mesh
    t = np.ascontiguousarray(mesh.t, dtype=np.int32)
    p = np.ascontiguousarray(mesh.p, dtype=np.float64)
    tdu = np.ascontiguousarray(asm.dofnum_u.t_dof, dtype=np.int32)
    sel = None if tind is None else np.ascontiguousarray(tind, dtype=np.int32)
    n = t.shape[1] if sel is None else sel.shape[0]
    out = np.zeros((plan.spec.n_entries, n))
    a = SpfemArgs()
    a.set_params(plan)
    a.n, a.ldo, a.ldk = n, n, 1
    a.sel = _ptr(sel) if sel is not None else None
    a.t, a.ldt = _ptr(t), t.shape[1]
    a.p, a.ldp = _ptr(p), p.shape[1]
    a.tdof_u, a.ldd = _ptr(tdu), tdu.shape[1]
    keep = []
    for s, key in enumerate(plan.wkeys):
        arr = np.ascontiguousarray(interp[key], dtype=np.float64)
        keep.append(arr)
        a.interp[s] = arr.ctypes.data
    for i, f in enumerate(getattr(plan, "fields", ()) or ()):
        arr = np.ascontiguousarray(f.T, dtype=np.float64)      # (nq, nt), mesh element order
        keep.append(arr)
        a.fields[i] = arr.ctypes.data
        a.ldf = arr.shape[1]
    a.out = _ptr(out)
    compile_host(plan)(C.byref(a))
    return out


def iasm(asm, form, intorder=None, tind=None, interp=None):
    """Host emulation of ``AssemblerElement.iasm`` (generated code + SciPy)."""
    plan = asm.plan_iasm(form, intorder=intorder, interp=interp)
    out = run_interior(asm, plan, tind, interp)
    sel = slice(None) if tind is None else np.asarray(tind)
    tv = asm.dofnum_v.t_dof[:, sel]
    n = tv.shape[1]
    if plan.bilinear:
        tu = asm.dofnum_u.t_dof[:, sel]
        nu, nv = tu.shape[0], tv.shape[0]
        rows = np.tile(tv, (nu, 1)).reshape(-1)
        cols = np.repeat(tu, nv, axis=0).reshape(-1)
        return coo_matrix((out.reshape(-1), (rows, cols)),
                          shape=(asm.dofnum_v.N, asm.dofnum_u.N)).tocsr()
    return coo_matrix((out.reshape(-1), (tv.reshape(-1), np.zeros(tv.size, dtype=np.int64))),
                      shape=(asm.dofnum_v.N, 1)).toarray().T[0]


def fasm(asm, form, find=None, interior=False, intorder=None, normals=True,
         interp=None):
    """Host emulation of ``AssemblerElement.fasm``."""
    mesh = asm.mesh
    if find is None:
        find = mesh.interior_facets() if interior else mesh.boundary_facets()
    plan = asm.plan_fasm(form, interior=interior, intorder=intorder,
                         normals=normals, interp=interp, find=find)
    fv, e1, e2, lf1 = asm.facet_tables(find)
    n = len(find)
    nblk = len(plan.spec.blocks)
    t = np.ascontiguousarray(mesh.t, dtype=np.int32)
    p = np.ascontiguousarray(mesh.p, dtype=np.float64)
    tdu = np.ascontiguousarray(asm.dofnum_u.t_dof, dtype=np.int32)
    fv = np.ascontiguousarray(fv, dtype=np.int32)
    e1 = np.ascontiguousarray(e1, dtype=np.int32)
    e2 = np.ascontiguousarray(e2, dtype=np.int32)
    lf1 = np.ascontiguousarray(lf1, dtype=np.int32)
    out = np.zeros((plan.spec.n_entries, nblk * n))
    a = SpfemArgs()
    a.set_params(plan)
    a.n, a.ldo, a.ldk = n, nblk * n, 1
    a.t, a.ldt = _ptr(t), t.shape[1]
    a.p, a.ldp = _ptr(p), p.shape[1]
    a.tdof_u, a.ldd = _ptr(tdu), tdu.shape[1]
    keep = []
    for s, key in enumerate(plan.wkeys):
        arr = np.ascontiguousarray(interp[key], dtype=np.float64)
        keep.append(arr)
        a.interp[s] = arr.ctypes.data
    a.fv, a.e1, a.e2, a.lf1 = _ptr(fv), _ptr(e1), _ptr(e2), _ptr(lf1)
    for i_, f in enumerate(getattr(plan, "fields", ()) or ()):
        arr = np.ascontiguousarray(f.T, dtype=np.float64)      # (nq, nfacets), order of `find`
        keep.append(arr)
        a.fields[i_] = arr.ctypes.data
        a.ldf = arr.shape[1]
    a.out = _ptr(out)
    compile_host(plan)(C.byref(a))
    rows, cols = asm.facet_dofs(find, interior)
    if plan.bilinear:
        nu, nv = cols.shape[0], rows.shape[0]
        R = np.tile(rows, (nu, 1)).reshape(-1)
        Cc = np.repeat(cols, nv, axis=0).reshape(-1)
        return coo_matrix((out.reshape(-1), (R, Cc)),
                          shape=(asm.dofnum_v.N, asm.dofnum_u.N)).tocsr()
    return coo_matrix((out.reshape(-1), (rows.reshape(-1), np.zeros(rows.size, dtype=np.int64))),
                      shape=(asm.dofnum_v.N, 1)).toarray().T[0]


def fan_iasm(asm, form, rows_per_tile=64, tiny=True):
    """Host emulation of the vertex-fan kernel: plan (tests/fanplan_ref.py) +
    the generated ``<name>_fan_host`` loop; pattern from the plain emulation."""
    from spfem_b200._lib import SpfemFanArgs
    from fanplan_ref import build_fan_plan, fill_coordinates_host
    ref = iasm(asm, form)
    plan = asm.plan_iasm(form)
    assert plan.spec.fan_ok
    fp = build_fan_plan(asm.mesh, ref.indptr, ref.indices, rows_per_tile, tiny=tiny)
    fill_coordinates_host(fp, asm.mesh.p)
    key = hashlib.sha256(plan.source.encode()).hexdigest()[:20]
    compile_host(plan)
    lib = C.CDLL(os.path.join(_TMP, key + ".so"))
    fn = getattr(lib, plan.name + "_fan_host")
    fn.argtypes = [C.POINTER(SpfemFanArgs), C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    fn.restype = None
    values = np.full(ref.nnz, np.nan)
    sV = np.zeros(fp["max_ns"] + 16)
    sG = np.zeros(fp["max_ns"] + 16, dtype=np.int32)
    a = SpfemFanArgs()
    a.base.set_params(plan)
    a.desc, a.blob, a.values = _ptr(fp["desc"]), _ptr(fp["blob"]), _ptr(values)
    a.max_a, a.max_b, a.max_ns = fp["max_a"], fp["max_b"], fp["max_ns"]
    fn(C.byref(a), fp["ntiles"], _ptr(sV), _ptr(sG),
       108 if fp["format"] == "tiny" else fp["max_neighbours"])
    out = ref.copy()
    out.data = values
    return out, fp


def inorm(asm, form, interp, intorder=None):
    """Host emulation of ``AssemblerElement.inorm`` (per-element values)."""
    captured = {}

    class _Eng(object):
        def run_functional(self, plan, interp_):
            captured["plan"] = plan
            return None
    saved = asm._engine
    asm._engine = _Eng()
    try:
        asm.inorm(form, interp, intorder=intorder)
    finally:
        asm._engine = saved
    return run_interior(asm, captured["plan"], None, interp)[0]


def fnorm(asm, form, interp, intorder=None, interior=False, normals=True):
    """Host emulation of ``AssemblerElement.fnorm`` (per-facet values)."""
    captured = {}

    class _Eng(object):
        def run_facet_functional(self, plan, find, interp_):
            captured["plan"], captured["find"] = plan, find
            return None
    saved = asm._engine
    asm._engine = _Eng()
    try:
        asm.fnorm(form, interp, intorder=intorder, interior=interior, normals=normals)
    finally:
        asm._engine = saved
    plan, find = captured["plan"], captured["find"]
    mesh = asm.mesh
    fv, e1, e2, lf1 = asm.facet_tables(find)
    n = len(find)
    t = np.ascontiguousarray(mesh.t, dtype=np.int32)
    p = np.ascontiguousarray(mesh.p, dtype=np.float64)
    tdu = np.ascontiguousarray(asm.dofnum_u.t_dof, dtype=np.int32)
    fv = np.ascontiguousarray(fv, dtype=np.int32)
    e1 = np.ascontiguousarray(e1, dtype=np.int32)
    e2 = np.ascontiguousarray(e2, dtype=np.int32)
    lf1 = np.ascontiguousarray(lf1, dtype=np.int32)
    out = np.zeros(n)
    a = SpfemArgs()
    a.set_params(plan)
    a.n, a.ldo, a.ldk = n, n, 1
    a.t, a.ldt = _ptr(t), t.shape[1]
    a.p, a.ldp = _ptr(p), p.shape[1]
    a.tdof_u, a.ldd = _ptr(tdu), tdu.shape[1]
    keep = []
    for s, key in enumerate(plan.wkeys):
        arr = np.ascontiguousarray(interp[key], dtype=np.float64)
        keep.append(arr)
        a.interp[s] = arr.ctypes.data
    a.fv, a.e1, a.e2, a.lf1 = _ptr(fv), _ptr(e1), _ptr(e2), _ptr(lf1)
    a.out = _ptr(out)
    compile_host(plan)(C.byref(a))
    return out, find
