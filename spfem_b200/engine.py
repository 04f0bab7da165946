# -*- coding: utf-8 -*-
"""Device side of :class:`spfem_b200.assembly.AssemblerElement`.

Owns the device copies of the mesh (PyTorch tensors: memory + streams only),
the compiled kernels and the cached sparsity patterns, and drives the C-ABI
library (``_lib.py``).  Layout in HBM:

* ``p``      float64 ``(dim, nv)`` row-major (SoA coordinates),
* ``t``      int32 ``(nve, nt)``  -- elements re-ordered along a Morton curve
  so that neighbouring threads touch neighbouring vertices / CSR rows; the
  permutation is internal: DOF numbering and therefore the CSR pattern are the
  reference's, only the (fp64) summation order inside a CSR slot changes,
* ``tdof_u/v`` int32 ``(Nbfun, nt)`` in the same element order,
* pattern: ``indptr, indices`` (int32, or int64 by SciPy's rule), ``cptr``
  (nnz+1) and ``contrib`` (Nu*Nv*nt) contribution lists,
* ``local``  float64 ``(Nu*Nv, nt)`` entry-major local matrices (scratch),
* ``values`` float64 ``(nnz,)``.
"""
import ctypes as C
import os
import time

import numpy as np
from scipy.sparse import csr_matrix

from . import _lib
from .codegen import source_key

_KCACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_kcache")
_MODULES = {}     # source hash -> (module handle, kernel handle, image bytes)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.SpfemError("spfem_b200 needs a CUDA device (B200, sm_100a); "
                              "there is no CPU fallback")
    return torch


def compile_cached(source):
    """cubin image of ``source`` (in-process + on-disk cache)."""
    key = source_key(source)
    path = os.path.join(_KCACHE_DIR, key + ".cubin")
    if os.path.exists(path):
        with open(path, "rb") as fh:
            return key, fh.read()
    image = _lib.nvrtc_compile(source)
    try:
        os.makedirs(_KCACHE_DIR, exist_ok=True)
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(image)
        os.replace(tmp, path)
    except OSError:
        pass
    return key, image


def get_kernel(source, name):
    key = source_key(source)
    hit = _MODULES.get(key)
    if hit is not None:
        return hit[1]
    lib = _lib.load()
    key, image = compile_cached(source)
    module = C.c_void_p()
    buf = C.create_string_buffer(image, len(image))
    _lib.check(lib.spfem_module_load(buf, len(image), C.byref(module)))
    kernel = C.c_void_p()
    _lib.check(lib.spfem_module_get_kernel(module, name.encode(), C.byref(kernel)))
    _MODULES[key] = (module, kernel, buf)
    return kernel


def scipy_index_dtype(n_entries, nrows, ncols):
    """Index width SciPy gives ``coo_matrix(...).tocsr()``: int32 unless the
    pre-deduplication COO length or a dimension exceeds 2^31-1
    (scipy/sparse/_coo.py, _compressed.py)."""
    return np.int64 if max(n_entries, nrows, ncols) > 2 ** 31 - 1 else np.int32


class Pattern(object):
    """CSR structure + contribution lists of one (row table, column table)."""

    def __init__(self, nrows, ncols, nnz, indptr, indices, cptr, contrib,
                 rowstart, n_entries):
        self.nrows, self.ncols, self.nnz = nrows, ncols, nnz
        self.indptr, self.indices = indptr, indices
        self.cptr, self.contrib, self.rowstart = cptr, contrib, rowstart
        self.n_entries = n_entries
        self._host = None

    def host(self):
        """Host copies of ``indptr`` / ``indices`` (cached)."""
        if self._host is None:
            self._host = (self.indptr.cpu().numpy(), self.indices.cpu().numpy())
        return self._host


class DeviceCSR(object):
    """Device-resident assembly result: the pattern + a ``values`` tensor."""

    def __init__(self, pattern, values, shape):
        self.pattern, self.values, self.shape = pattern, values, shape

    @property
    def nnz(self):
        return self.pattern.nnz

    def to_scipy(self):
        indptr, indices = self.pattern.host()
        data = self.values.cpu().numpy()
        A = csr_matrix(self.shape, dtype=np.float64)
        A.data = data
        A.indices = indices.copy()
        A.indptr = indptr.copy()
        A.has_sorted_indices = True
        A.has_canonical_format = True
        return A


class Engine(object):
    def __init__(self, asm, device=None, morton=True):
        self.torch = torch = _torch()
        self.lib = _lib.load()
        self.asm = asm
        self.device = torch.device("cuda", torch.cuda.current_device()
                                   if device is None else device)
        self.patterns = {}
        self.stats = {}
        mesh = asm.mesh
        self.nt = mesh.t.shape[1]
        self.nv = mesh.p.shape[1]
        if max(self.nv, asm.dofnum_u.N, asm.dofnum_v.N) > 2 ** 31 - 1:
            raise NotImplementedError("more than 2^31-1 vertices / DOFs")
        self._morton = morton and mesh.refdom in ("tri", "tet", "quad")
        self.upload_mesh(first=True)

    # -- helpers -----------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, arr, dtype):
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype))
        return t.to(self.device, non_blocking=False)

    def _empty(self, n, dtype):
        return self.torch.empty(int(n), dtype=dtype, device=self.device)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def upload_mesh(self, first=False):
        """(Re-)copy the host mesh arrays ``p`` and ``t`` to the device.  The
        connectivity-derived tables (Morton order, DOF tables, patterns) are
        built once; later calls only refresh coordinates and connectivity
        values (same topology), which is what a time-stepping / moving-mesh
        loop -- and the end-to-end benchmark -- needs."""
        torch = self.torch
        mesh, asm = self.asm.mesh, self.asm
        self.p = self._dev(mesh.p, np.float64)
        t32 = self._dev(mesh.t, np.int32)
        self.h2d_bytes = self.p.numel() * 8 + t32.numel() * 4
        if first:
            if self._morton and self.nt > 1:
                self.perm = self._morton_perm(t32)
            else:
                self.perm = None
            self.invperm = None
            tdu = self._dev(asm.dofnum_u.t_dof, np.int32)
            self.tdof_u = tdu if self.perm is None else tdu.index_select(1, self.perm_l)
            if asm.dofnum_v is asm.dofnum_u:
                self.tdof_v = self.tdof_u
            else:
                tdv = self._dev(asm.dofnum_v.t_dof, np.int32)
                self.tdof_v = tdv if self.perm is None else tdv.index_select(1, self.perm_l)
        self.t = t32 if self.perm is None else t32.index_select(1, self.perm_l).contiguous()

    def _morton_perm(self, t32):
        torch = self.torch
        mesh = self.asm.mesh
        dim = mesh.p.shape[0]
        nve, nt = t32.shape
        perm = self._empty(nt, torch.int32)
        wsb = self.lib.spfem_morton_ws_bytes(nt)
        ws = self._empty(wsb, torch.uint8)
        lo = (C.c_double * 3)(*([float(mesh.p[d].min()) for d in range(dim)] + [0.0] * (3 - dim)))
        hi = (C.c_double * 3)(*([float(mesh.p[d].max()) for d in range(dim)] + [1.0] * (3 - dim)))
        _lib.check(self.lib.spfem_morton_order(
            self._ptr(self.p), self.p.shape[1], dim, self._ptr(t32), nt, nve, nt,
            lo, hi, self._ptr(perm), self._ptr(ws), wsb, self._stream()))
        self.perm_l = perm.long()
        return perm

    def _inverse_perm(self):
        if self.invperm is None and self.perm is not None:
            torch = self.torch
            inv = torch.empty(self.nt, dtype=torch.int64, device=self.device)
            inv[self.perm_l] = torch.arange(self.nt, device=self.device)
            self.invperm = inv
        return self.invperm

    def _elements(self, ids):
        """Positions (in the device element order) of the host element ids."""
        torch = self.torch
        ids = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).to(self.device)
        if self.perm is None:
            return ids
        neg = ids < 0
        out = self._inverse_perm()[ids.clamp(min=0)]
        out[neg] = -1
        return out

    # -- pattern -----------------------------------------------------------
    def build_pattern(self, rowdof, coldof, nrows, ncols):
        """``rowdof`` (Nv, n) / ``coldof`` (Nu, n) int32 device tensors (or
        ``coldof=None`` for a linear form)."""
        torch, lib = self.torch, self.lib
        t0 = time.perf_counter()
        nloc_v, n = rowdof.shape
        nloc_u = coldof.shape[0] if coldof is not None else 1
        n_entries = n * nloc_v * nloc_u
        rowdof = rowdof.contiguous()
        if coldof is not None:
            coldof = coldof.contiguous()
        wsb = lib.spfem_pattern_ws_bytes(n_entries)
        ws = self._empty(wsb, torch.uint8)
        nnz = C.c_longlong(0)
        _lib.check(lib.spfem_pattern_sort(
            self._ptr(rowdof), n, self._ptr(coldof), n, n, nloc_v, nloc_u,
            nrows, ncols, self._ptr(ws), wsb, C.byref(nnz), self._stream()))
        nnz = nnz.value
        contrib = self._empty(max(n_entries, 1), torch.int32)
        if coldof is None:
            rowstart = self._empty(nrows + 1, torch.int32)
            _lib.check(lib.spfem_pattern_fill(
                self._ptr(ws), n_entries, nrows, ncols, nnz, 4, None, None,
                None, self._ptr(contrib), self._ptr(rowstart), self._stream()))
            pat = Pattern(nrows, ncols, nnz, None, None, None, contrib,
                          rowstart, n_entries)
        else:
            idt = scipy_index_dtype(n_entries, nrows, ncols)
            tdt = torch.int64 if idt == np.int64 else torch.int32
            indptr = self._empty(nrows + 1, tdt)
            indices = self._empty(max(nnz, 1), tdt)[:nnz]
            cptr = self._empty(nnz + 1, torch.int32)
            _lib.check(lib.spfem_pattern_fill(
                self._ptr(ws), n_entries, nrows, ncols, nnz,
                8 if idt == np.int64 else 4, self._ptr(indptr),
                self._ptr(indices), self._ptr(cptr), self._ptr(contrib), None,
                self._stream()))
            pat = Pattern(nrows, ncols, nnz, indptr, indices, cptr, contrib,
                          None, n_entries)
        torch.cuda.synchronize(self.device)
        del ws
        self.stats["pattern_s"] = time.perf_counter() - t0
        return pat

    # -- launches ----------------------------------------------------------
    def _interp_tensors(self, plan, interp):
        out = []
        for key in plan.wkeys:
            out.append(self._dev(np.asarray(interp[key]), np.float64))
        return out

    def _reduce(self, local, pat, bilinear):
        torch, lib = self.torch, self.lib
        if bilinear:
            values = self._empty(max(pat.nnz, 1), torch.float64)[:pat.nnz]
            _lib.check(lib.spfem_reduce(self._ptr(local), self._ptr(pat.cptr),
                                        self._ptr(pat.contrib), pat.nnz,
                                        self._ptr(values), self._stream()))
        else:
            values = self._empty(pat.nrows, torch.float64)
            _lib.check(lib.spfem_reduce(self._ptr(local), self._ptr(pat.rowstart),
                                        self._ptr(pat.contrib), pat.nrows,
                                        self._ptr(values), self._stream()))
        return values

    def run_interior(self, plan, tind=None, interp=None, to_host=True):
        torch, lib, asm = self.torch, self.lib, self.asm
        spec = plan.spec
        kernel = get_kernel(plan.source, plan.name)
        if tind is None:
            n, sel = self.nt, None
            key = ("interior", plan.bilinear)
            pat = self.patterns.get(key)
            if pat is None:
                pat = self.build_pattern(self.tdof_v, self.tdof_u if plan.bilinear else None,
                                         asm.dofnum_v.N, asm.dofnum_u.N if plan.bilinear else 1)
                self.patterns[key] = pat
        else:
            tind = np.asarray(tind)
            if tind.dtype == bool:
                tind = np.nonzero(tind)[0]
            pos = self._elements(tind)
            n = int(pos.shape[0])
            sel = pos.to(torch.int32)
            rowdof = self.tdof_v.index_select(1, pos)
            coldof = self.tdof_u.index_select(1, pos) if plan.bilinear else None
            pat = self.build_pattern(rowdof, coldof, asm.dofnum_v.N,
                                     asm.dofnum_u.N if plan.bilinear else 1)
        local = self._empty(max(spec.n_entries * n, 1), torch.float64)
        wt = self._interp_tensors(plan, interp)
        a = _lib.SpfemArgs()
        a.n, a.ldo = n, n
        a.sel = sel.data_ptr() if sel is not None else None
        a.t, a.ldt = self.t.data_ptr(), self.t.shape[1]
        a.p, a.ldp = self.p.data_ptr(), self.p.shape[1]
        a.tdof_u, a.ldd = self.tdof_u.data_ptr(), self.tdof_u.shape[1]
        for s, w in enumerate(wt):
            a.interp[s] = w.data_ptr()
        a.out = local.data_ptr()
        _lib.check(lib.spfem_launch(kernel, C.byref(a), 128, self._stream()))
        values = self._reduce(local, pat, plan.bilinear)
        return self._finish(values, pat, plan.bilinear, to_host)

    def run_facet(self, plan, find, interior, interp=None, to_host=True):
        torch, lib, asm = self.torch, self.lib, self.asm
        spec = plan.spec
        kernel = get_kernel(plan.source, plan.name)
        find = np.asarray(find)
        if find.dtype == bool:
            find = np.nonzero(find)[0]
        n = int(find.shape[0])
        nblk = len(spec.blocks)
        fv, e1, e2, lf1 = asm.facet_tables(find)
        rows, cols = asm.facet_dofs(find, interior)
        rowdof = self._dev(rows, np.int32)
        coldof = self._dev(cols, np.int32) if plan.bilinear else None
        pat = self.build_pattern(rowdof, coldof, asm.dofnum_v.N,
                                 asm.dofnum_u.N if plan.bilinear else 1)
        fv_d = self._dev(fv, np.int32)
        e1_d = self._elements(e1).to(torch.int32)
        e2_d = self._elements(e2).to(torch.int32)
        lf_d = self._dev(lf1, np.int32)
        local = self._empty(max(spec.n_entries * nblk * n, 1), torch.float64)
        wt = self._interp_tensors(plan, interp)
        a = _lib.SpfemArgs()
        a.n, a.ldo = n, nblk * n
        a.t, a.ldt = self.t.data_ptr(), self.t.shape[1]
        a.p, a.ldp = self.p.data_ptr(), self.p.shape[1]
        a.tdof_u, a.ldd = self.tdof_u.data_ptr(), self.tdof_u.shape[1]
        for s, w in enumerate(wt):
            a.interp[s] = w.data_ptr()
        a.fv, a.e1, a.e2, a.lf1 = (fv_d.data_ptr(), e1_d.data_ptr(),
                                   e2_d.data_ptr(), lf_d.data_ptr())
        a.out = local.data_ptr()
        _lib.check(lib.spfem_launch(kernel, C.byref(a), 64, self._stream()))
        values = self._reduce(local, pat, plan.bilinear)
        return self._finish(values, pat, plan.bilinear, to_host)

    def _finish(self, values, pat, bilinear, to_host):
        asm = self.asm
        if bilinear:
            res = DeviceCSR(pat, values, (asm.dofnum_v.N, asm.dofnum_u.N))
            return res.to_scipy() if to_host else res
        return values.cpu().numpy() if to_host else values
