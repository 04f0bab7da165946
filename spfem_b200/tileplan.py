# -*- coding: utf-8 -*-
"""Tile plan of the fused assembly kernel (``codegen.TILE_KERNEL``): Python side
of the native builder ``spfem_tileplan_build`` (csrc/spfem_plan.cu, declared in
include/spfem_b200.h).

The elements (already in Morton order) are cut into tiles of ``E`` consecutive
elements.  Every CSR row is *owned* by the tile that contains most of its
elements (ties: the lowest tile; ``owner_rule = 0``: simply the lowest tile); a
tile's CTA computes the local matrices of all elements that touch a
row it owns (its own elements plus a thin halo that neighbouring tiles compute
as well) into shared memory and then produces every CSR value of its rows
exactly once, summing the contributions in a fixed order.  No floating point
atomics and no partial sums in HBM: the result is bitwise reproducible.

The builder also produces the CSR structure itself (``indptr`` / ``indices`` of
SciPy's ``coo_matrix(...).tocsr()``, spfem/assembly.py:985-986), tile batch by
tile batch, so that nothing of the size ``Nu*Nv*nt`` ever exists at once; one
batch = one launch of the kernel.  This module only owns the memory (PyTorch
tensors handed to the library through allocator callbacks) and offers NumPy
emulations of what the kernel reads, for the tests without a GPU.

Blob of one tile (all sections 16-byte aligned; descriptor = 16 int32):

  section A (TMA copy 1, needed to compute local matrices)
    vt     uint16 [nve][ne_pad]   TILE-LOCAL vertex ids of the tile's elements
    el     int32  [ne_pad]        element ids (optional; interp gathers only)
    pc     f64    2-D: [nvt_pad][2] interleaved, 3-D: [3][nvt_pad]  coordinates of
                                  the tile's vertices (private copy, refreshed by
                                  ``fill_coordinates``)
    sbase  int32 [ne_pad], smask uint32 [W][ne_pad]   (compact plans only: offset
                                  of the element's packed values + bit mask of the
                                  distinct local values that are read at all)
  section B (TMA copy 2, needed to sum and store)
    pbase  int64           lowest CSR position of the tile's rows (+ 8 bytes pad)
    slots ordered by contribution count (``sort_slots != 0``):
      spos uint32 [ns]     CSR position of the slot's (first) value minus pbase
      slen uint16 [ns]     block plans only: length of the slot's scalar row (the
                           bsv sub-rows of a block are bsu * length apart)
    slots in CSR order (``sort_slots = 0``; the tile's rows one after the other):
      rrec {uint32 pos, uint16 first, uint16 len} [nr]   per ROW: position of its
                           first value minus pbase, its first slot, its length
      shdr {uint32 mask, uint32 row0} [ceil(ns/32)]      per 32 slots: bit l set =
                           slot l starts a new row; row of slot 0.  The row of slot
                           l is row0 + popcount(mask & bits 0..l)
    soff   uint16 [ns+1]   start of every slot's contributions in ``contrib``
    contrib uint16 [nc+1]  scalar plans: shared-memory cell ``u*ldl + el`` (compact:
                           packed position); block plans: ``(j*nbv+i) << els | el``
  The slots of a tile are in CSR order (``sort_slots = 0``) or ordered by their
  number of contributions (1: exactly, 2: coarse classes, CSR order inside a class).

Descriptor: 0 blob offset / 16, 1 bytes A, 2 bytes B, 3 ne, 4 ne_pad, 5 ns,
6 / 7 / 8 byte offsets of soff / slen (or shdr) / contrib in B, 9 nr, 10 nc, 11 nvt_pad,
12 / 13 byte offsets of pc / el in A (13 = -1: no element ids), 14 nvt,
15 byte offset of sbase in A.
"""
import ctypes as C

import numpy as np

from . import _lib


class TileTooLarge(ValueError):
    """A tile does not fit the kernel's 16-bit fields: retry with a smaller E."""


def _pad16(x):
    return (x + 15) // 16 * 16


class _Arena(object):
    """Allocator callbacks of the native builder: every buffer is a PyTorch
    tensor (device memory stays owned by PyTorch); ``out:*`` tensors are kept."""

    def __init__(self, torch, device):
        self.torch, self.device = torch, device
        self.live = {}
        self.kept = {}
        self.peak = self.cur = 0
        self.error = None

        def alloc(_ctx, nbytes, tag):
            try:
                t = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
            except Exception as exc:       # out of memory: reported by the library
                self.error = exc
                return None
            ptr = (t.data_ptr() + 255) // 256 * 256
            self.live[ptr] = (t, ptr - t.data_ptr(), int(nbytes), tag.decode())
            self.cur += int(nbytes)
            self.peak = max(self.peak, self.cur)
            return ptr

        def free(_ctx, ptr):
            hit = self.live.pop(ptr, None)
            if hit is not None:
                self.cur -= hit[2]

        self.alloc_cb = _lib.ALLOC_FN(alloc)
        self.free_cb = _lib.FREE_FN(free)

    def tensor(self, ptr, dtype, count=None):
        """Typed view of the live allocation that starts at ``ptr``."""
        t, off, nbytes, _tag = self.live[ptr]
        v = t[off:off + nbytes].view(dtype)
        return v if count is None else v[:count]


def build_tile_plan(torch, rowdof, coldof, vt, nrows, ncols, E, alias, nuniq,
                    dim, bsv=1, bsu=1, with_el=False, sort_slots=0, compact=False,
                    row_mask=None, index_bytes=4, batch_cap=0, split_factor=0.0,
                    ne_budget=0, owner_rule=1, stream=None):
    """Pattern + tile plan for the DOF tables ``rowdof`` (nloc_v, n) /
    ``coldof`` (nloc_u, n) or ``None`` (int32, one device, element order of the
    kernel) and the vertex table ``vt`` (nve, n).  Returns a dict; raises
    :class:`TileTooLarge` when ``E`` has to shrink."""
    dev = rowdof.device
    on_dev = dev.type == "cuda"
    lib = _lib.load()
    rowdof = rowdof.contiguous()
    vt = vt.contiguous()
    if coldof is not None:
        coldof = coldof.contiguous()
    al = torch.as_tensor(np.ascontiguousarray(alias, dtype=np.int32)).to(dev)
    pin = _lib.SpfemPlanIn()
    pin.rowdof, pin.ldr, pin.nloc_v = rowdof.data_ptr(), rowdof.shape[1], rowdof.shape[0]
    if coldof is not None:
        pin.coldof, pin.ldc, pin.nloc_u = coldof.data_ptr(), coldof.shape[1], coldof.shape[0]
    else:
        pin.coldof, pin.ldc, pin.nloc_u = None, 0, 1
    pin.n = rowdof.shape[1]
    pin.nrows, pin.ncols = int(nrows), int(ncols)
    pin.vt, pin.ldvt, pin.nve, pin.dim = vt.data_ptr(), vt.shape[1], vt.shape[0], int(dim)
    pin.bsv, pin.bsu = int(bsv), int(bsu)
    pin.alias, pin.nuniq = al.data_ptr(), int(nuniq)
    pin.E, pin.with_el = int(E), 1 if with_el else 0
    pin.sort_slots = int(sort_slots)
    pin.compact = 1 if compact else 0
    if row_mask is not None:
        row_mask = row_mask.to(torch.uint8).contiguous()
        pin.row_mask = row_mask.data_ptr()
    pin.index_bytes = int(index_bytes)
    pin.batch_cap = int(batch_cap)
    pin.split_factor = float(split_factor)
    pin.ne_budget = int(ne_budget)
    pin.owner_rule = int(owner_rule)
    pout = _lib.SpfemPlanOut()
    arena = _Arena(torch, dev)
    if stream is None and on_dev:
        stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.spfem_tileplan_build(C.byref(pin), C.byref(pout), arena.alloc_cb, arena.free_cb,
                                  None, 1 if on_dev else 0, C.c_void_p(stream or 0))
    if rc != 0:
        msg = lib.spfem_last_error().decode("utf-8", "replace")
        arena.live.clear()
        if rc == 2:
            raise TileTooLarge(msg)
        if rc == 3 and arena.error is not None:
            raise arena.error
        raise _lib.SpfemError(msg)
    idt = torch.int64 if index_bytes == 8 else torch.int32
    plan = {
        "nnz": int(pout.nnz), "ldl": int(pout.ldl), "els": int(pout.els), "E": int(E),
        "bsv": int(pout.bsv), "bsu": int(pout.bsu), "nve": vt.shape[0], "dim": int(dim),
        "ntiles": int(pout.ntiles), "compact": bool(compact), "sort_slots": int(sort_slots),
        "with_el": bool(with_el), "nuniq": int(nuniq),
        "max_ne": int(pout.max_ne), "max_nr": int(pout.max_nr), "max_ns": int(pout.max_ns),
        "max_store": int(pout.max_store), "max_bytes_a": int(pout.max_bytes_a),
        "max_bytes_b": int(pout.max_bytes_b),
        "halo_factor": float(pout.n_pairs) / float(pin.n),
        "n_contrib": int(pout.n_contrib),
        "indptr": arena.tensor(pout.indptr, idt, int(nrows) + 1),
        "indices": (arena.tensor(pout.indices, idt, int(pout.nnz))
                    if pout.indices else None),
        "scratch_peak_bytes": arena.peak,
        "batches": [],
    }
    blob_bytes = 0
    for b in range(pout.nbatches):
        B = pout.batches[b]
        nt = int(B.ntiles)
        plan["batches"].append({
            "ntiles": nt, "first_tile": int(B.first_tile), "nslots": int(B.nslots),
            "desc": arena.tensor(B.desc, torch.int32, 16 * nt),
            "blob": arena.tensor(B.blob, torch.uint8, int(B.blob_bytes)),
            "vid": arena.tensor(B.vid, torch.int32, int(B.nvid)),
            "vptr": arena.tensor(B.vptr, torch.int64, nt + 1),
        })
        blob_bytes += int(B.blob_bytes)
    plan["blob_bytes"] = blob_bytes
    arena.live.clear()
    return plan


def fill_coordinates(plan, p, stream=None):
    """Copy the vertex coordinates ``p`` (dim, nv) into the tiles' private
    coordinate sections (call after every change of ``p``)."""
    lib = _lib.load()
    on_dev = p.device.type == "cuda"
    if stream is None and on_dev:
        import torch
        stream = torch.cuda.current_stream(p.device).cuda_stream
    p = p.contiguous()
    for B in plan["batches"]:
        _lib.check(lib.spfem_tileplan_fill_coords(
            B["desc"].data_ptr(), B["blob"].data_ptr(), B["vid"].data_ptr(),
            B["vptr"].data_ptr(), B["ntiles"], p.data_ptr(), p.shape[1], plan["dim"],
            1 if on_dev else 0, C.c_void_p(stream or 0)))


# ---------------------------------------------------------------------------
# NumPy emulations of what the kernel reads (tests only)
# ---------------------------------------------------------------------------
def _sections(plan, B, d, blob):
    off16, ba, bb, ne, ne_pad, ns, o_soff, o_srow, o_cb, nr, nc = [int(x) for x in d[:11]]
    A = blob[off16 * 16: off16 * 16 + ba]
    S = blob[off16 * 16 + ba: off16 * 16 + ba + bb]
    return A, S, ne, ne_pad, ns, o_soff, o_srow, o_cb, nr, nc


def emulate(plan, local, ut=None):
    """Phase B of the tile kernel in NumPy: ``local`` is the ``(nuniq, n)`` array
    of the elements' distinct local values (device element order); ``ut`` the
    ``(nbv*nbu, bsv*bsu)`` table of distinct-value ids of a block plan.  Needs a
    plan built ``with_el=True``.  Reads only the blobs, like the kernel.
    Returns ``(values, written)``."""
    nnz = plan["nnz"] if plan["indices"] is not None else int(plan["indptr"].numel() - 1)
    ldl, els = plan["ldl"], plan["els"]
    bsv, bsu = plan["bsv"], plan["bsu"]
    nb = bsv * bsu
    values = np.full(nnz, np.nan)
    written = np.zeros(nnz, dtype=np.int64)
    nuniq = local.shape[0]
    for B in plan["batches"]:
        desc = B["desc"].cpu().numpy().astype(np.int64).reshape(-1, 16)
        blob = B["blob"].cpu().numpy()
        for d in desc:
            A, S, ne, ne_pad, ns, o_soff, o_srow, o_cb, nr, nc = _sections(plan, B, d, blob)
            if ne == 0 or ns == 0:
                continue
            assert d[13] >= 0, "emulate() needs with_el=True"
            elems = A[d[13]: d[13] + 4 * ne].view(np.int32).astype(np.int64)
            if plan["compact"]:
                W = (nuniq + 31) // 32
                ob = int(d[15])
                sb = A[ob: ob + 4 * ne_pad].view(np.int32).astype(np.int64)
                sm = A[ob + 4 * ne_pad: ob + 4 * (1 + W) * ne_pad].view(np.uint32).reshape(W, ne_pad)
                sL = np.full(max(plan["max_store"], 1), np.nan)
                for el in range(ne):
                    pos = sb[el]
                    for u in range(nuniq):
                        if (int(sm[u >> 5, el]) >> (u & 31)) & 1:
                            sL[pos] = local[u, elems[el]]
                            pos += 1
            else:
                sL = np.full(nuniq * ldl + 16, np.nan)
                for u in range(nuniq):
                    sL[u * ldl: u * ldl + ne] = local[u, elems]
            pbase = int(S[:8].view(np.int64)[0])
            if plan["sort_slots"]:
                spos = S[16: 16 + 4 * ns].view(np.uint32).astype(np.int64)
                slen = (S[o_srow: o_srow + 2 * ns].view(np.uint16).astype(np.int64)
                        if nb > 1 else np.zeros(ns, dtype=np.int64))
            else:
                rrec = S[16: 16 + 8 * nr].view(np.uint32).astype(np.int64).reshape(nr, 2)
                nch = (ns + 31) // 32
                shdr = S[o_srow: o_srow + 8 * nch].view(np.uint32).astype(np.int64).reshape(nch, 2)
                spos, slen = np.zeros(ns, dtype=np.int64), np.zeros(ns, dtype=np.int64)
                for s in range(ns):
                    mask, row0 = shdr[s >> 5]
                    row = row0 + bin(int(mask) & (0xffffffff >> (31 - (s & 31)))).count("1")
                    pos, fl = rrec[row]
                    spos[s] = pos + bsu * (s - (fl & 0xffff))
                    slen[s] = fl >> 16
            soff = S[o_soff: o_soff + 2 * (ns + 1)].view(np.uint16).astype(np.int64)
            cb = S[o_cb: o_cb + 2 * (nc + 1)].view(np.uint16).astype(np.int64)
            assert soff[ns] == nc
            for s in range(ns):
                g = pbase + spos[s]
                L = slen[s]
                acc = np.zeros(nb)
                for kk in range(soff[s], soff[s + 1]):
                    c = cb[kk]
                    if nb == 1:
                        acc[0] += sL[c]
                    else:
                        el, sidx = c & ((1 << els) - 1), c >> els
                        for ab in range(nb):
                            acc[ab] += sL[int(ut[sidx, ab]) * ldl + el]
                for a in range(bsv):
                    for b in range(bsu):
                        pos = g + a * bsu * L + b
                        values[pos] = acc[a * bsu + b]
                        written[pos] += 1
    return values, written


def check_vertex_table(plan, vt, p):
    """Tests only: section A of every blob holds, through the tile-local vertex
    ids, the coordinates of the vertices of the tile's elements."""
    nve, dim = plan["nve"], plan["dim"]
    vt = np.asarray(vt)
    for B in plan["batches"]:
        desc = B["desc"].cpu().numpy().astype(np.int64).reshape(-1, 16)
        blob = B["blob"].cpu().numpy()
        for d in desc:
            off16, ba, ne, ne_pad, nvt_pad, off_pc = d[0], d[1], d[3], d[4], d[11], d[12]
            A = blob[off16 * 16: off16 * 16 + ba]
            loc = A[:nve * ne_pad * 2].view(np.uint16).reshape(nve, ne_pad)[:, :ne]
            pc = A[off_pc: off_pc + dim * nvt_pad * 8].view(np.float64)
            pc = pc.reshape(nvt_pad, 2).T if dim == 2 else pc.reshape(dim, nvt_pad)
            elems = A[d[13]: d[13] + 4 * ne].view(np.int32).astype(np.int64)
            for r in range(nve):
                for k in range(dim):
                    if not np.array_equal(pc[k, loc[r]], p[k, vt[r, elems]]):
                        return False
    return True
