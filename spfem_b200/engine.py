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


def tile_store(nuniq, tp):
    """Doubles of shared memory reserved for a tile's local values."""
    if tp["compact"]:
        return tp["max_store"] + 2
    return nuniq * tp["ldl"]


def tile_smem_bytes(nuniq, tp, nsidx=0, nb=1):
    """Dynamic shared memory of the tile kernel (layout in codegen.TILE_KERNEL)."""
    sl = (8 * tile_store(nuniq, tp) + 15) // 16 * 16
    ut = 4 * nsidx * nb if nb > 1 else 0
    return sl + tp["max_bytes_a"] + tp["max_bytes_b"] + 16 + ut + 16


def codegen_tile_threads():
    from . import codegen
    return codegen.TILE_THREADS

_KCACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_kcache")
_ARCH = "sm_100a"
_MODULES = {}     # cache key -> (module handle, {kernel name: handle}, image buffer); LRU
_MAX_MODULES = int(os.environ.get("SPFEM_MAX_MODULES", "64"))
_MAX_KCACHE_FILES = int(os.environ.get("SPFEM_MAX_KCACHE_FILES", "512"))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.SpfemError("spfem_b200 needs a CUDA device (B200, sm_100a); "
                              "there is no CPU fallback")
    return torch


def _cache_key(source):
    """Key of a compiled kernel: the source AND what it was compiled for / with
    (the same text gives a different cubin on another architecture or NVRTC)."""
    return source_key("%s|abi%d|%s" % (_ARCH, 4, source))


def _trim_kcache():
    """Keep the on-disk cubin cache bounded (oldest files out)."""
    try:
        files = [os.path.join(_KCACHE_DIR, f) for f in os.listdir(_KCACHE_DIR)
                 if f.endswith(".cubin")]
        if len(files) <= _MAX_KCACHE_FILES:
            return
        files.sort(key=os.path.getmtime)
        for f in files[:len(files) - _MAX_KCACHE_FILES]:
            os.remove(f)
    except OSError:
        pass


def compile_cached(source):
    """cubin image of ``source`` (in-process + on-disk cache)."""
    key = _cache_key(source)
    path = os.path.join(_KCACHE_DIR, key + ".cubin")
    if os.path.exists(path):
        with open(path, "rb") as fh:
            image = fh.read()
        try:
            os.utime(path)                      # most recently used
        except OSError:
            pass
        return key, image
    image = _lib.nvrtc_compile(source, _ARCH)
    try:
        os.makedirs(_KCACHE_DIR, exist_ok=True)
        tmp = path + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(image)
        os.replace(tmp, path)
        _trim_kcache()
    except OSError:
        pass
    return key, image


def get_kernel(source, name):
    """Kernel handle ``name`` of the module generated from ``source``.  Loaded
    modules are kept in a bounded LRU; the evicted ones are unloaded
    (``spfem_module_unload``), so a long-running process that keeps tracing new
    forms does not accumulate device code."""
    key = _cache_key(source)
    hit = _MODULES.pop(key, None)
    lib = _lib.load()
    if hit is None:
        key, image = compile_cached(source)
        module = C.c_void_p()
        buf = C.create_string_buffer(image, len(image))
        _lib.check(lib.spfem_module_load(buf, len(image), C.byref(module)))
        hit = (module, {}, buf)
    _MODULES[key] = hit                         # (re-)insert as most recently used
    if len(_MODULES) > _MAX_MODULES:
        # evict the least recently used modules that no staged assembly holds
        for k in [k for k in _MODULES if _PINS.get(k, 0) == 0][:len(_MODULES) - _MAX_MODULES]:
            if k == key:
                continue
            old = _MODULES.pop(k)
            try:
                _torch().cuda.synchronize()     # its kernels may still be queued
                lib.spfem_module_unload(old[0])
            except Exception:
                pass
    kernels = hit[1]
    if name not in kernels:
        kernel = C.c_void_p()
        _lib.check(lib.spfem_module_get_kernel(hit[0], name.encode(), C.byref(kernel)))
        kernels[name] = kernel
    return kernels[name]


_PINS = {}        # cache key -> number of live Prepared objects using the module


def pin_module(source, owner):
    """Keep the module of ``source`` loaded while ``owner`` is alive."""
    import weakref
    key = _cache_key(source)
    _PINS[key] = _PINS.get(key, 0) + 1

    def unpin(k=key):
        _PINS[k] = _PINS.get(k, 1) - 1
        if _PINS[k] <= 0:
            _PINS.pop(k, None)
    weakref.finalize(owner, unpin)


def scipy_index_dtype(n_entries, nrows, ncols):
    """Index width SciPy gives ``coo_matrix(...).tocsr()``: int32 unless the
    pre-deduplication COO length or a dimension exceeds 2^31-1
    (scipy/sparse/_coo.py, _compressed.py)."""
    return np.int64 if max(n_entries, nrows, ncols) > 2 ** 31 - 1 else np.int32


class Pattern(object):
    """CSR structure (+ the contribution lists of the two-kernel route; ``None``
    when the pattern came out of the native tile-plan builder)."""

    def __init__(self, nrows, ncols, nnz, indptr, indices, cptr, contrib,
                 rowstart, n_entries):
        self.nrows, self.ncols, self.nnz = nrows, ncols, nnz
        self.indptr, self.indices = indptr, indices
        self.cptr, self.contrib, self.rowstart = cptr, contrib, rowstart
        self.n_entries = n_entries
        self._host = None

    def host(self):
        """Host copies of ``indptr`` / ``indices`` (cached, read-only)."""
        if self._host is None:
            ip, ix = self.indptr.cpu().numpy(), self.indices.cpu().numpy()
            ip.flags.writeable = False
            ix.flags.writeable = False
            self._host = (ip, ix)
        return self._host


class DeviceCSR(object):
    """Device-resident assembly result: the pattern + a ``values`` tensor."""

    def __init__(self, pattern, values, shape):
        self.pattern, self.values, self.shape = pattern, values, shape

    @property
    def nnz(self):
        return self.pattern.nnz

    def to_scipy(self):
        """Host ``scipy.sparse.csr_matrix`` with fresh arrays, like the
        reference's: the values are copied asynchronously at PCIe speed into a
        recycled pinned block while worker threads copy the cached host pattern.

        ``SPFEM_SHARE_PATTERN=1`` (opt-in, for time loops that assemble on the
        same mesh again and again): ``indptr`` / ``indices`` are then the cached
        arrays themselves, shared by every result of this assembler and
        READ-ONLY -- 8 bytes per nonzero per call instead of 12-16, and N ranks
        on one host do not contend for memory bandwidth copying identical index
        arrays.  SciPy operations that build new matrices work as usual;
        in-place structure changes (``eliminate_zeros``, ``sort_indices`` on an
        unsorted matrix) raise on the read-only arrays, which is why it is not
        the default."""
        indptr, indices = self.pattern.host()
        A = csr_matrix(self.shape, dtype=np.float64)
        pending = _to_host_begin(self.values)       # DMA runs while we copy the pattern
        if os.environ.get("SPFEM_SHARE_PATTERN", "0") == "1":
            A.indices, A.indptr = indices, indptr
        else:
            A.indices = _host_copy(indices, pooled=True)
            A.indptr = _host_copy(indptr, pooled=True)
        A.data = _to_host_finish(pending)
        A.has_sorted_indices = True
        A.has_canonical_format = True
        return A


_POOL = None


def _chunks(n, parts):
    step = (n + parts - 1) // parts
    return [(i, min(n, i + step)) for i in range(0, n, step)]


def _parallel(fn, n, min_chunk=1 << 21):
    """Run ``fn(lo, hi)`` over [0, n) on a few worker threads (NumPy releases
    the GIL inside copies / casts)."""
    global _POOL
    parts = min(8, max(1, n // min_chunk))
    if parts <= 1:
        fn(0, n)
        return
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=8)
    list(_POOL.map(lambda ab: fn(*ab), _chunks(n, parts)))


def _host_copy(arr, pooled=False):
    """Fresh copy of a cached host array (threaded).  ``pooled``: into a buffer
    of the recycling pool -- a loop that drops its previous result gets the same
    pages back, without the first-touch page faults of a new allocation (they,
    not the copy, dominated: 4.9 ms -> 1-2 ms for the 117 MB of config 2)."""
    out = None
    if pooled and arr.size > (1 << 20) and arr.ndim == 1:
        import torch
        tdt = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}.get(arr.dtype)
        if tdt is not None:
            out, _buf = _PINNED.take(tdt, arr.size)
    if out is None:
        out = np.empty_like(arr)
    if arr.size == 0:
        return out

    def work(lo, hi):
        out[lo:hi] = arr[lo:hi]
    _parallel(work, arr.shape[0])
    return out


class _PinnedPool(object):
    """Recycling pool of pinned host buffers for results that are handed to the
    caller.  ``take`` returns a NumPy array backed by a pinned tensor; when the
    array is garbage collected its buffer goes back to the pool (weakref
    finalizer), so a loop that drops the previous result re-uses the same pinned
    block: device-to-host copies run at PCIe speed straight into the memory the
    caller owns, with no staging copy and no repeated page-locking."""

    def __init__(self, max_cached_bytes=2 << 30, max_held_bytes=None):
        self.free = {}            # (dtype, capacity) -> [tensor, ...]
        self.cached = 0
        self.max_cached = max_cached_bytes
        # page-locked bytes the callers still hold (results they keep alive): above
        # the budget new results come in pageable memory (staged copy), so that a
        # user who keeps many matrices does not exhaust lockable memory
        self.held = 0
        if max_held_bytes is None:
            max_held_bytes = int(os.environ.get("SPFEM_PINNED_BUDGET", str(16 << 30)))
        self.max_held = max_held_bytes

    def take(self, torch_dtype, n):
        import torch
        import weakref
        cap = max(int(n), 1)
        if n > (1 << 20):
            cap = -(-int(n) // (1 << 20)) * (1 << 20)        # round up to 1 Mi elements
        key = (torch_dtype, cap)
        lst = self.free.get(key)
        if lst:
            buf = lst.pop()
            self.cached -= buf.numel() * buf.element_size()
        else:
            nbytes = cap * torch.empty(0, dtype=torch_dtype).element_size()
            pin = self.held + nbytes <= self.max_held
            buf = torch.empty(cap, dtype=torch_dtype, pin_memory=pin)
        self.held += buf.numel() * buf.element_size()
        # the finalizer hangs on the array that wraps the whole buffer: every view
        # the caller (or SciPy) derives keeps that array alive through ``.base``
        full = buf.numpy()
        weakref.finalize(full, self._release, key, buf)
        return full[:n], buf

    def _release(self, key, buf):
        nbytes = buf.numel() * buf.element_size()
        self.held -= nbytes
        if buf.is_pinned() and self.cached + nbytes <= self.max_cached:
            self.free.setdefault(key, []).append(buf)
            self.cached += nbytes


_PINNED = _PinnedPool()


def _to_host_begin(t, nchunks=8):
    """Start the asynchronous device-to-host copy of a 1-D tensor into a pinned
    buffer from the recycling pool; returns a handle for :func:`_to_host_finish`."""
    import torch
    n = t.numel()
    if n == 0:
        return np.empty(0, dtype=torch.empty(0, dtype=t.dtype).numpy().dtype), None
    arr, buf = _PINNED.take(t.dtype, n)
    buf[:n].copy_(t, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return arr, ev


def _to_host_finish(handle):
    arr, ev = handle
    if ev is not None:
        ev.synchronize()
    return arr


def _to_host(t):
    """1-D device tensor -> fresh NumPy array (see :func:`_to_host_begin`)."""
    return _to_host_finish(_to_host_begin(t))


class Prepared(object):
    """A ready-to-launch assembly: ``launch()`` enqueues the generated element
    kernel and the fixed-order reduction on the current stream and returns the
    (device) values tensor -- no host synchronisation, no allocation."""

    def __init__(self, engine, plan, kernel, args, block, local, pattern, keep=()):
        self.engine, self.plan, self.kernel = engine, plan, kernel
        self.args, self.block, self.local, self.pattern = args, block, local, pattern
        self._keep = keep
        pin_module(plan.source, self)
        torch = engine.torch
        n = pattern.nnz if plan.bilinear else pattern.nrows
        self.values = engine._empty(max(n, 1), torch.float64)[:n]
        self.n_launches = 2

    def launch_element_kernel(self):
        e = self.engine
        a = self.args
        a.t, a.ldt = e.t.data_ptr(), e.t.shape[1]
        a.p, a.ldp = e.p.data_ptr(), e.p.shape[1]
        _lib.check(e.lib.spfem_launch(self.kernel, C.byref(a), self.block, e._stream()))

    def launch_reduce(self):
        e, pat = self.engine, self.pattern
        if self.plan.bilinear:
            cptr, nslots = pat.cptr, pat.nnz
        else:
            cptr, nslots = pat.rowstart, pat.nrows
        _lib.check(e.lib.spfem_reduce(e._ptr(self.local), e._ptr(cptr),
                                      e._ptr(pat.contrib), nslots,
                                      e._ptr(self.values), e._stream()))

    def launch(self):
        self.launch_element_kernel()
        self.launch_reduce()
        return self.values

    def result(self, to_host=True):
        """The assembled matrix / vector of the last ``launch()``."""
        return self.engine._finish(self.values, self.pattern,
                                   self.plan.bilinear, to_host)


class TiledPrepared(Prepared):
    """Fused variant: ``<name>_tile`` (codegen.TILE_KERNEL) once per batch of
    the tile plan -- local matrices stay in shared memory and every CSR value is
    written once."""

    def __init__(self, engine, plan, kernel, tile_kernel, args, pattern, tp, keep=()):
        Prepared.__init__(self, engine, plan, kernel, args, 128, None, pattern, keep)
        self.tile_kernel, self.tp = tile_kernel, tp
        self._kernel_version = None
        if tile_kernel is None:
            self._select_kernel()
        spec = plan.spec
        nb = tp["bsv"] * tp["bsu"]
        nsidx = (spec.n_rows // tp["bsv"]) * (spec.n_cols // tp["bsu"])
        self.smem = tile_smem_bytes(spec.n_unique, tp, nsidx, nb)
        self.targs = []
        for B in tp["batches"]:
            ta = _lib.SpfemTileArgs()
            ta.desc = B["desc"].data_ptr()
            ta.blob = B["blob"].data_ptr()
            ta.values = self.values.data_ptr()
            ta.ldl = tp["ldl"]
            ta.max_a = tp["max_bytes_a"]
            ta.max_b = tp["max_bytes_b"]
            ta.compact = 1 if tp["compact"] else 0
            ta.store = tile_store(spec.n_unique, tp)
            ta.sorted = 1 if tp["sort_slots"] else 0
            ta.els = tp["els"]
            ta.prefetch = int(os.environ.get("SPFEM_TILE_PREFETCH", "300"))   # about one wave ahead
            self.targs.append((ta, B["ntiles"]))
        # 256 threads measured best (128 and 512 are 0-40 % slower)
        self.threads = min(256, codegen_tile_threads())
        if os.environ.get("SPFEM_TILE_THREADS"):
            self.threads = min(int(os.environ["SPFEM_TILE_THREADS"]), codegen_tile_threads())
        self.n_launches = len(self.targs)

    def _select_kernel(self):
        """``<name>_tile`` of the plan's module -- for quadrilateral forms with a
        parallelogram branch (``spec.affine_variant``) the variant compiled without
        the quadrature loop while every element of the mesh is a parallelogram
        (re-checked when the coordinates change)."""
        e, plan = self.engine, self.plan
        source = plan.source
        if getattr(plan.spec, "affine_variant", False) and e.all_parallelograms():
            source = "#define SPFEM_PARALLELOGRAMS 1\n" + source
            pin_module(source, self)
        self.tile_kernel = get_kernel(source, plan.name + "_tile")
        self._kernel_version = e.p_version

    def launch(self):
        e = self.engine
        if self._kernel_version is not None and self._kernel_version != e.p_version:
            self._select_kernel()
        a = self.args
        a.t, a.ldt = e.t.data_ptr(), e.t.shape[1]
        a.p, a.ldp = e.p.data_ptr(), e.p.shape[1]
        st = e._stream()
        for ta, ntiles in self.targs:
            ta.base = a
            _lib.check(e.lib.spfem_launch_tiled(self.tile_kernel, C.byref(ta), ntiles,
                                                self.threads, self.smem, st))
        return self.values


class FanPrepared(Prepared):
    """Vertex-fan variant (scalar P1 on triangles, element-constant
    coefficients): one launch of ``<name>_fan`` (codegen.FAN_KERNEL), one thread
    per CSR row, local rows recomputed from coordinates in registers."""

    def __init__(self, engine, plan, kernel, fan_kernel, args, pattern, fp):
        Prepared.__init__(self, engine, plan, kernel, args, 128, None, pattern, ())
        self.fan_kernel, self.fp = fan_kernel, fp
        fa = _lib.SpfemFanArgs()
        fa.desc = fp["desc"].data_ptr()
        fa.blob = fp["blob"].data_ptr()
        fa.values = self.values.data_ptr()
        fa.max_a, fa.max_b, fa.max_ns = fp["max_a"], fp["max_b"], fp["max_ns"]
        # L2 prefetch distance in tiles (about two waves of resident CTAs ahead).
        # (Measured and rejected in round 2 as well: several tiles per CTA with
        # double-buffered TMA copies -- 2 ... 32 tiles per CTA are 6 ... 46 % slower,
        # profiles/r02_fan_tiles_per_cta.jsonl.)
        fa.prefetch = int(os.environ.get("SPFEM_FAN_PREFETCH", "2400"))
        self.fargs = fa
        self.smem = fp["max_a"] + 12 * fp["max_ns"] + 64
        from . import codegen
        self.threads = getattr(plan.spec, "fan_threads", codegen.FAN_THREADS)
        self.n_launches = 1
        self.tp = {"E": fp["rows_per_tile"], "ldl": 0, "ntiles": fp["ntiles"],
                   "halo_factor": fp["halo_factor"], "kind": "vertex-fan"}

    def launch(self):
        e = self.engine
        self.fargs.base = self.args
        _lib.check(e.lib.spfem_launch_tiled(self.fan_kernel, C.byref(self.fargs),
                                            self.fp["ntiles"], self.threads,
                                            self.smem, e._stream()))
        return self.values


class Engine(object):
    def __init__(self, asm, device=None, morton=True):
        self.torch = torch = _torch()
        self.lib = _lib.load()
        self.asm = asm
        self.device = torch.device("cuda", torch.cuda.current_device()
                                   if device is None else device)
        self.patterns = {}
        self.stats = {}
        self._tile_plans = []
        self._tile_cache = {}
        self._facet_cache = {}        # facet set -> (pattern, index tables)
        self.p_version = 0
        self.fused = os.environ.get("SPFEM_FUSED", "1") != "0"
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
        if first:
            # pinned staging buffers: host arrays are converted/copied into
            # them and sent with asynchronous copies at PCIe speed
            self._pin_p = torch.empty(mesh.p.shape, dtype=torch.float64, pin_memory=True)
            self._pin_t = torch.empty(mesh.t.shape, dtype=torch.int32, pin_memory=True)
            self._h2d_done = None
        if self._h2d_done is not None:
            self._h2d_done.synchronize()         # the staging buffers are free again
        pp, pt = self._pin_p.numpy(), self._pin_t.numpy()
        hp, ht = mesh.p, mesh.t

        def cp_p(lo, hi):
            pp[:, lo:hi] = hp[:, lo:hi]

        def cp_t(lo, hi):
            pt[:, lo:hi] = ht[:, lo:hi]      # int64 -> int32 on the fly
        # (finer chunks pipelined with the DMA were measured slower: the host side,
        # not the 168 MB over PCIe, is the longer part)
        _parallel(cp_p, hp.shape[1], 1 << 19)
        self.p = self._pin_p.to(self.device, non_blocking=True)
        _parallel(cp_t, ht.shape[1], 1 << 19)     # staged while p is on the wire
        t32 = self._pin_t.to(self.device, non_blocking=True)
        self._h2d_done = torch.cuda.Event()
        self._h2d_done.record()
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
        if not first:
            # the fused kernel reads tile-private copies of the coordinates
            from .tileplan import fill_coordinates
            from .fanplan import fill_fan_coordinates
            self.p_version += 1
            for tp in self._tile_plans:
                if tp.get("kind") == "vertex-fan":
                    fill_fan_coordinates(tp, self.p)
                else:
                    fill_coordinates(tp, self.p)
                tp["p_version"] = self.p_version

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

    def _field_tensors(self, plan, permute=True):
        """Device copies ``(nq, n)`` of the plan's host-evaluated coefficient
        fields (interior forms: in the device element order; facet forms: in the
        order of the facet set)."""
        torch = self.torch
        out = []
        for f in getattr(plan, "fields", ()) or ():
            t = torch.from_numpy(np.ascontiguousarray(f.T)).to(self.device)
            if permute and self.perm is not None:
                t = t.index_select(1, self.perm_l).contiguous()
            out.append(t)
        return out

    @staticmethod
    def _set_fields(a, ft):
        for i, t in enumerate(ft):
            a.fields[i] = t.data_ptr()
        a.ldf = ft[0].shape[1] if ft else 0

    def _sorted_pattern(self, bilinear):
        """Whole-mesh pattern with contribution lists (global radix sort): what
        the two-kernel route and the vertex-fan plan need."""
        asm = self.asm
        key = ("interior", bilinear)
        pat = self.patterns.get(key)
        if pat is None:
            pat = self.build_pattern(self.tdof_v, self.tdof_u if bilinear else None,
                                     asm.dofnum_v.N, asm.dofnum_u.N if bilinear else 1)
            self.patterns[key] = pat
        return pat

    def prepare_interior(self, plan, tind=None, interp=None):
        """Everything up to the launch: compiled kernel, pattern, plan, output
        tensors, argument block.  Returns a :class:`Prepared` whose ``launch()``
        only enqueues kernels on the current stream.  Whole-mesh assemblies take
        the fused routes (vertex fan for scalar P1 on triangles, element tiles
        otherwise); ``tind`` subsets and forms whose halo recompute would be too
        expensive take the two-kernel route."""
        torch, asm = self.torch, self.asm
        spec = plan.spec
        kernel = get_kernel(plan.source, plan.name)
        wt = self._interp_tensors(plan, interp)
        ft = self._field_tensors(plan)
        wt = (wt, ft)                      # kept alive by the Prepared object
        a = _lib.SpfemArgs()
        a.set_params(plan)
        a.tdof_u, a.ldd = self.tdof_u.data_ptr(), self.tdof_u.shape[1]
        for s, w in enumerate(wt[0]):
            a.interp[s] = w.data_ptr()
        self._set_fields(a, ft)
        if tind is None:
            n, sel = self.nt, None
            a.n, a.ldo, a.ldk = n, n, 1
            if self.fused and n > 0:
                if (plan.bilinear and getattr(spec, "fan_ok", False)
                        and os.environ.get("SPFEM_FAN", "1") != "0"
                        and spec.n_entries * n <= 2 ** 31 - 1):
                    pat = self._sorted_pattern(True)
                    fp = self.fan_plan(pat, getattr(spec, "fan_rows", 128))
                    if fp is not None:
                        fk = get_kernel(plan.source, plan.name + {
                            "tiny": "_fan8t", "fan8": "_fan8"}.get(fp["format"], "_fan"))
                        return FanPrepared(self, plan, kernel, fk, a, pat, fp)
                hit = self.tile_plan(plan)
                if hit is not None:
                    pat, tp = hit
                    return TiledPrepared(self, plan, kernel, None, a, pat, tp, keep=(wt,))
            if spec.n_entries * n > 2 ** 31 - 1:
                raise NotImplementedError(
                    "this form needs the two-kernel route, which is limited to 2^31-1 "
                    "local entries per call (%d elements x %d)" % (n, spec.n_entries))
            pat = self._sorted_pattern(plan.bilinear)
        else:
            tind = np.asarray(tind)
            if tind.dtype == bool:
                tind = np.nonzero(tind)[0]
            pos = self._elements(tind)
            n = int(pos.shape[0])
            a.n, a.ldo, a.ldk = n, n, 1
            sel = pos.to(torch.int32)
            rowdof = self.tdof_v.index_select(1, pos)
            coldof = self.tdof_u.index_select(1, pos) if plan.bilinear else None
            pat = self.build_pattern(rowdof, coldof, asm.dofnum_v.N,
                                     asm.dofnum_u.N if plan.bilinear else 1)
        a.sel = sel.data_ptr() if sel is not None else None
        local = self._empty(max(spec.n_entries * n, 1), torch.float64)
        a.out = local.data_ptr()
        return Prepared(self, plan, kernel, a, 128, local, pat, keep=(sel, wt))

    def all_parallelograms(self):
        """Every element of a quadrilateral mesh is a parallelogram -- the exact test
        of the generated kernels (``codegen.looped_quad_body``), once per coordinate
        upload."""
        if getattr(self, "_para_version", None) != self.p_version:
            self._para = False
            if self.t.shape[0] == 4 and self.p.shape[0] == 2 and \
                    os.environ.get("SPFEM_QPARA", "1") != "0":
                t, p = self.t.long(), self.p
                d = p[:, t[0]] - p[:, t[1]] + p[:, t[2]] - p[:, t[3]]
                self._para = bool((d == 0).all())
            self._para_version = self.p_version
        return self._para

    def fan_plan(self, pat, rows=128):
        """Plan of the vertex-fan kernel for ``pat`` with ``rows`` rows per tile
        (cached on the pattern), or ``None`` when the assembler is not
        scalar-P1-on-triangles shaped."""
        cache = pat.__dict__.setdefault("_fan_plans", {})
        if rows in cache:
            return cache[rows]
        asm, torch = self.asm, self.torch
        cache[rows] = None
        mesh = asm.mesh
        if (asm.dofnum_u is not asm.dofnum_v or mesh.refdom != "tri"
                or asm.dofnum_u.t_dof.shape != mesh.t.shape
                or not np.array_equal(asm.dofnum_u.t_dof, mesh.t)):
            return None
        from .fanplan import build_fan_plan, FanPlanError, fill_fan_coordinates
        t0 = time.perf_counter()
        try:
            # (the device connectivity is in Morton order: any element order gives a
            # valid plan, only the start of a closed ring may differ)
            fp = build_fan_plan(torch, self.p, self.t, pat.indptr, pat.indices, int(rows),
                                tiny=os.environ.get("SPFEM_FAN_TINY", "1") != "0",
                                row_order=os.environ.get("SPFEM_FAN_ROWORDER", "id"))
        except FanPlanError:
            return None
        fill_fan_coordinates(fp, self.p)
        fp["kind"] = "vertex-fan"
        self._tile_plans.append(fp)
        self.torch.cuda.synchronize(self.device)
        self.stats["fan_plan_s"] = time.perf_counter() - t0
        cache[rows] = fp
        return fp

    def _row_mask(self):
        rm = getattr(self.asm, "row_mask", None)
        if rm is None:
            return None
        if getattr(self, "_row_mask_t", None) is None:
            self._row_mask_t = self._dev(np.asarray(rm), np.uint8)
        return self._row_mask_t

    def tile_plan(self, plan):
        """``(pattern, tile plan)`` of the fused kernel for the whole mesh (built
        natively, ``tileplan.build_tile_plan``; cached per distinct-value layout),
        or ``None`` when the fused kernel does not apply."""
        from .tileplan import build_tile_plan, fill_coordinates, TileTooLarge
        torch, asm, spec = self.torch, self.asm, plan.spec
        env = os.environ.get
        ckey = (plan.bilinear, spec.n_unique, tuple(spec.alias), spec.needs_el,
                env("SPFEM_TILE_MODE", "auto"), env("SPFEM_TILE_E", "0"),
                env("SPFEM_TILE_SORT", "auto"), env("SPFEM_TILE_BLOCKS", "1"),
                env("SPFEM_TILE_SPLIT", "0"), env("SPFEM_TILE_BUDGET", "1"),
                env("SPFEM_TILE_OWNER", "1"))
        if ckey in self._tile_cache:
            return self._tile_cache[ckey]
        force = env("SPFEM_FUSED", "1") == "force"
        n_terms = getattr(spec, "n_terms", 0)
        dim = self.p.shape[0]
        # the fused kernel pays off while recomputing the halo elements is cheap
        # (multiply-adds per element x halo factor; per-point geometry of the Q2
        # element -> two-kernel route, profiles/r01_configs.json)
        if not force and n_terms * (1.3 if dim == 2 else 2.5) > 1500:
            self._tile_cache[ckey] = None
            return None
        bsv, bsu = spec.bsv, spec.bsu
        if env("SPFEM_TILE_BLOCKS", "1") == "0":
            bsv = bsu = 1
        nb = bsv * bsu
        if nb > 1 and (bsv != spec.bsv or bsu != spec.bsu):
            nb = 1
        mode = env("SPFEM_TILE_MODE", "auto")
        # 3-D scalar meshes: the halo is 2-3 x the tile, keep only the values that
        # are read (measured: TetP2 17 %, TetP1 12 % faster); otherwise dense image
        compact = (dim == 3 and nb == 1) if mode == "auto" else mode == "compact"
        if nb > 1:
            compact = False
        srt = env("SPFEM_TILE_SORT", "auto")
        if srt == "auto":
            # 3-D: a row's slots have 1 ... ~24 contributions; ordering a tile's
            # slots by that count keeps warps converged (TetP2 +25 %, TetP1 +14 %)
            # 2-D vector elements: coarse count classes in CSR order (+6 %)
            sort_slots = 1 if dim == 3 else (2 if nb > 1 else 0)
        else:
            sort_slots = {"bucket": 2, "1": 1}.get(srt, 0)
        nuniq = spec.n_unique
        nsidx = (spec.n_rows // spec.bsv) * (spec.n_cols // spec.bsu)
        minb = int(env("SPFEM_TILE_MINB", str(getattr(spec, "minb", 3 if dim == 2 else 2))))
        if getattr(spec, "affine_variant", False) and self.all_parallelograms():
            minb = int(env("SPFEM_TILE_MINB_AFFINE", str(spec.minb_affine)))
        # two CTAs per SM (3-D, Q2), else four for scalar plans (TriP2: 48 registers) and
        # three for the block plans of vector elements (80 registers)
        prefer = (110 if minb <= 2 or compact else 58 if nb > 1 else 55) * 1024
        limit = 220 * 1024
        E_env = int(env("SPFEM_TILE_E", "0"))
        if E_env > 0:
            cands = [E_env]
            while cands[-1] > 16:
                cands.append(max(16, cands[-1] * 3 // 4 // 16 * 16))
        else:
            # measured (profiles/r01_configs.json): three to four CTAs per SM
            # (<= 56 KB each) but at least 64 elements per tile
            cands = [384, 256, 160, 128, 96, 72, 64, 48, 32, 16]
        bilinear = plan.bilinear
        n_entries = spec.n_entries * self.nt
        idt = scipy_index_dtype(n_entries, asm.dofnum_v.N, asm.dofnum_u.N if bilinear else 1)
        t0 = time.perf_counter()

        def build(E, ne_budget=0):
            return build_tile_plan(torch, self.tdof_v, self.tdof_u if bilinear else None,
                                   self.t, asm.dofnum_v.N, asm.dofnum_u.N if bilinear else 1,
                                   E, spec.alias, nuniq, dim,
                                   bsv=bsv if nb > 1 else 1, bsu=bsu if nb > 1 else 1,
                                   with_el=spec.needs_el, sort_slots=sort_slots, compact=compact,
                                   row_mask=self._row_mask(),
                                   index_bytes=8 if idt == np.int64 else 4,
                                   batch_cap=int(env("SPFEM_TILE_BATCH", "0")),
                                   split_factor=float(env("SPFEM_TILE_SPLIT", "0")),
                                   ne_budget=ne_budget,
                                   owner_rule=int(env("SPFEM_TILE_OWNER", "1")))
        tp = None
        for E in cands:
            if E_env <= 0 and not compact and E > 64 and 8 * nuniq * (1.25 * E + 24) > prefer:
                continue                # cannot fit the preferred footprint
            if E_env <= 0 and compact and E > 72:
                continue
            # quadrature-loop bodies are long: one pass of phase A over the CTA's threads
            # (NPART threads per tile element) beats a second, half-empty pass
            if (E_env <= 0 and getattr(spec, "looped", False) and spec.n_unique > 24 and E > 64
                    and 1.25 * E * spec.npart > codegen_tile_threads()):
                continue
            try:
                tp = build(E)
            except TileTooLarge:
                tp = None
                continue
            if nb > 1 and (tp["bsv"], tp["bsu"]) != (bsv, bsu):
                tp = None               # DOF tables without the block structure
                break
            smem = tile_smem_bytes(nuniq, tp, nsidx, nb)
            if smem > limit:
                tp = None
                continue
            if E_env > 0 or smem <= prefer or E <= 64:
                break
            tp = None
        # The largest tile sizes the shared memory of every CTA.  The Morton curve
        # does not follow the cells of a grid whose size is not a power of two, so a
        # tail of tiles has far more halo elements than the typical one: if splitting
        # that tail lets one more CTA share an SM, rebuild with an element budget.
        if tp is not None and env("SPFEM_TILE_BUDGET", "1") != "0":
            smem = tile_smem_bytes(nuniq, tp, nsidx, nb)
            per_sm = 227 * 1024
            cmax = min(minb, 3 if dim == 2 else 2)        # CTAs per SM the registers allow
            c_now = min(cmax, per_sm // (smem + 1024))
            mean_ne = tp["halo_factor"] * self.nt / max(tp["ntiles"], 1)
            fixed = tp["max_bytes_b"] + 64 + (4 * nsidx * nb if nb > 1 else 0)
            per_ne = ((8.0 * (tp["max_store"] + 2) if compact else 8.0 * nuniq * tp["ldl"])
                      + tp["max_bytes_a"]) / max(tp["max_ne"], 1)
            for c in range(cmax, c_now, -1):
                budget = int((per_sm // c - 1024 - fixed) / per_ne) - 2
                if budget < 1.04 * mean_ne or budget >= tp["max_ne"]:
                    continue
                try:
                    tp2 = build(tp["E"], ne_budget=budget)
                except TileTooLarge:
                    break
                smem2 = tile_smem_bytes(nuniq, tp2, nsidx, nb)
                if min(cmax, per_sm // (smem2 + 1024)) > c_now:
                    tp = tp2
                break
        if tp is not None and not force and n_terms * tp["halo_factor"] > 1500:
            tp = None
        if tp is not None and not bilinear and tp["nnz"] != asm.dofnum_v.N:
            tp = None                   # untouched / masked DOFs: the dense vector needs zeros there
        hit = None
        if tp is not None:
            fill_coordinates(tp, self.p)
            tp["p_version"] = self.p_version
            tp["kind"] = "element-tile"
            self._tile_plans.append(tp)
            pkey = ("tile", bilinear)
            pat = self.patterns.get(pkey)
            if pat is None:
                nnz = tp["nnz"] if bilinear else asm.dofnum_v.N
                pat = Pattern(asm.dofnum_v.N, asm.dofnum_u.N if bilinear else 1, nnz,
                              tp["indptr"] if bilinear else None, tp["indices"],
                              None, None, None, n_entries)
                self.patterns[pkey] = pat
            tp["indptr"] = tp["indices"] = None       # one copy, owned by the pattern
            hit = (pat, tp)
        self.torch.cuda.synchronize(self.device)
        self.stats["tile_plan_s"] = time.perf_counter() - t0
        self._tile_cache[ckey] = hit
        return hit

    def run_interior(self, plan, tind=None, interp=None, to_host=True):
        prep = self.prepare_interior(plan, tind, interp)
        values = prep.launch()
        return self._finish(values, prep.pattern, plan.bilinear, to_host)

    def run_functional(self, plan, interp):
        """One value per element (``inorm``): the generated kernel writes a single
        entry per element; the result is returned in the mesh's element order."""
        torch, lib = self.torch, self.lib
        kernel = get_kernel(plan.source, plan.name)
        n = self.nt
        out = self._empty(max(n, 1), torch.float64)
        wt = self._interp_tensors(plan, interp)
        a = _lib.SpfemArgs()
        a.set_params(plan)
        a.n, a.ldo, a.ldk = n, n, 1
        a.t, a.ldt = self.t.data_ptr(), self.t.shape[1]
        a.p, a.ldp = self.p.data_ptr(), self.p.shape[1]
        a.tdof_u, a.ldd = self.tdof_u.data_ptr(), self.tdof_u.shape[1]
        for s, w in enumerate(wt):
            a.interp[s] = w.data_ptr()
        a.out = out.data_ptr()
        _lib.check(lib.spfem_launch(kernel, C.byref(a), 128, self._stream()))
        if self.perm is not None:
            res = torch.empty_like(out)
            res[self.perm_l] = out
            out = res
        return _to_host(out)

    def prepare_facet(self, plan, find, interior, interp=None):
        """Facet assembly up to the launch.  The index tables and the sparsity
        pattern of a facet set are cached per (facet set, interior, bilinear,
        element pair): a time loop that calls ``fasm`` with the same facets again
        only traces, launches and copies (bounded cache, least recently used out)."""
        import hashlib
        torch, asm = self.torch, self.asm
        spec = plan.spec
        kernel = get_kernel(plan.source, plan.name)
        find = np.asarray(find)
        if find.dtype == bool:
            find = np.nonzero(find)[0]
        find = np.ascontiguousarray(find, dtype=np.int64)
        n = int(find.shape[0])
        nblk = len(spec.blocks)
        key = (bool(interior), bool(plan.bilinear), n,
               hashlib.sha1(find.tobytes()).hexdigest())
        hit = self._facet_cache.pop(key, None)
        if hit is None:
            fv, e1, e2, lf1 = asm.facet_tables(find)
            rows, cols = asm.facet_dofs(find, interior)
            rowdof = self._dev(rows, np.int32)
            coldof = self._dev(cols, np.int32) if plan.bilinear else None
            pat = self.build_pattern(rowdof, coldof, asm.dofnum_v.N,
                                     asm.dofnum_u.N if plan.bilinear else 1)
            hit = (pat, self._dev(fv, np.int32), self._elements(e1).to(torch.int32),
                   self._elements(e2).to(torch.int32), self._dev(lf1, np.int32))
        self._facet_cache[key] = hit                  # (re-)insert as most recent
        while len(self._facet_cache) > 8:
            self._facet_cache.pop(next(iter(self._facet_cache)))
        pat, fv_d, e1_d, e2_d, lf_d = hit
        local = self._empty(max(spec.n_entries * nblk * n, 1), torch.float64)
        wt = self._interp_tensors(plan, interp)
        ft = self._field_tensors(plan, permute=False)
        wt = (wt, ft)
        a = _lib.SpfemArgs()
        a.set_params(plan)
        self._set_fields(a, ft)
        a.n, a.ldo, a.ldk = n, nblk * n, 1
        a.tdof_u, a.ldd = self.tdof_u.data_ptr(), self.tdof_u.shape[1]
        for s, w in enumerate(wt[0]):
            a.interp[s] = w.data_ptr()
        a.fv, a.e1, a.e2, a.lf1 = (fv_d.data_ptr(), e1_d.data_ptr(),
                                   e2_d.data_ptr(), lf_d.data_ptr())
        a.out = local.data_ptr()
        return Prepared(self, plan, kernel, a, 64, local, pat,
                        keep=(fv_d, e1_d, e2_d, lf_d, wt))

    def run_facet(self, plan, find, interior, interp=None, to_host=True):
        prep = self.prepare_facet(plan, find, interior, interp)
        values = prep.launch()
        return self._finish(values, prep.pattern, plan.bilinear, to_host)

    def run_facet_functional(self, plan, find, interp):
        """One value per facet (``fnorm``): the generated facet kernel has a
        single accumulator; no pattern, no reduction."""
        torch, asm = self.torch, self.asm
        kernel = get_kernel(plan.source, plan.name)
        n = int(find.shape[0])
        fv, e1, e2, lf1 = asm.facet_tables(find)
        fv_d = self._dev(fv, np.int32)
        e1_d = self._elements(e1).to(torch.int32)
        e2_d = self._elements(e2).to(torch.int32)
        lf_d = self._dev(lf1, np.int32)
        out = self._empty(max(n, 1), torch.float64)
        wt = self._interp_tensors(plan, interp)
        a = _lib.SpfemArgs()
        a.set_params(plan)
        a.n, a.ldo, a.ldk = n, n, 1
        a.t, a.ldt = self.t.data_ptr(), self.t.shape[1]
        a.p, a.ldp = self.p.data_ptr(), self.p.shape[1]
        a.tdof_u, a.ldd = self.tdof_u.data_ptr(), self.tdof_u.shape[1]
        for s, w in enumerate(wt):
            a.interp[s] = w.data_ptr()
        a.fv, a.e1, a.e2, a.lf1 = (fv_d.data_ptr(), e1_d.data_ptr(),
                                   e2_d.data_ptr(), lf_d.data_ptr())
        a.out = out.data_ptr()
        _lib.check(self.lib.spfem_launch(kernel, C.byref(a), 64, self._stream()))
        return _to_host(out[:n])

    def _finish(self, values, pat, bilinear, to_host):
        asm = self.asm
        if bilinear:
            res = DeviceCSR(pat, values, (asm.dofnum_v.N, asm.dofnum_u.N))
            return res.to_scipy() if to_host else res
        return _to_host(values) if to_host else values
