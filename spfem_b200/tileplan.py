# -*- coding: utf-8 -*-
"""Tile plan of the fused assembly kernel (``codegen.TILE_KERNEL``).

The elements (already in Morton order) are cut into tiles of ``E`` consecutive
elements.  Every CSR row is *owned* by the lowest tile that contains one of its
elements; a tile's CTA computes the local matrices of all elements that touch a
row it owns (its own elements plus a thin halo that neighbouring tiles compute
as well) into shared memory and then produces every CSR value of its rows
exactly once, summing the contributions in a fixed order.  No floating point
atomics and no partial sums in HBM: the result is bitwise reproducible.

Built once per sparsity pattern from the global contribution lists
(``cptr``/``contrib`` of ``spfem_pattern_fill``) with PyTorch index operations
on the device (set-up plumbing; the tensors work on the CPU too, which is how
``tests/test_tileplan.py`` checks the plan without a GPU).

Everything a CTA needs besides the vertex coordinates is packed into ONE
contiguous, 16-byte aligned byte blob per tile so that it can be brought into
shared memory with two TMA bulk copies (``cp.async.bulk``):

  section A (needed to compute local matrices)
    vt     uint16 [nve][ne_pad]   TILE-LOCAL vertex ids of the tile's elements
    el     int32  [ne_pad]        element ids (optional; interp gathers only)
    pc     f64    [dim][nvt_pad]  coordinates of the tile's vertices (a private
                                  copy, refreshed by ``fill_coordinates`` whenever
                                  the mesh coordinates are uploaded)
  section B (needed to sum and store)
    rinfo  int32x2 [nr+1] per owned row: {CSR position of the row's first slot
                         minus its tile slot index, tile slot index of the first
                         slot | start of the row's contribution list << 16}
    contrib uint16 [nc]  shared-memory index ``idx*ldl + el`` of a contribution;
                         bit 15 marks the last contribution of a CSR slot.  Every
                         row's list is re-ordered by ``spfem_schedule_rows`` so
                         that the 16 rows of a half-warp never hit the same
                         shared-memory bank in the same step (idle steps read a
                         zero cell); slots of a row may finish in any order
    perm   uint8  [ns]   CSR position (inside its row) of the j-th finished slot

Descriptor (int32 x 16 per tile):
  0 blob offset / 16   1 bytes A   2 bytes B   3 ne   4 ne_pad   5 ns
  6 byte offset of perm in B   7 unused   8 byte offset of contrib in B   9 nr  10 nc
  11 nvt_pad  12 byte offset of pc in A  13 byte offset of el in A  14 nvt  15 0
"""
import numpy as np


def _pad(x, m):
    return (x + m - 1) // m * m


def _wrap32(torch, x):
    """Unsigned 32-bit values stored through an int32 view."""
    return torch.where(x >= 2 ** 31, x - 2 ** 32, x).to(torch.int32)


def _wrap16(torch, x):
    """Unsigned 16-bit values stored through an int16 view."""
    return torch.where(x >= 32768, x - 65536, x).to(torch.int16)


def build_tile_plan(torch, indptr, cptr, contrib, n, nent, E, vt, with_el=True, dim=2,
                    alias=None, schedule="inslot", mode="rows", owner_rule="lowest",
                    sort_slots=False):
    """All tensor arguments live on one device.  ``vt``: int32 ``(nve, n)``
    vertex table in device element order.  Returns a dict of tensors/ints.

    ``mode="rows"``: section B as described above (one thread per owned row).
    ``mode="slots"``: one thread per CSR slot (large local matrices: many short
    independent sums instead of few long serial ones); section B is then

      gdelta int32  [nr]     CSR position of the row's first slot minus its tile
                             slot index (tile slots are in CSR order)
      soff   uint16 [ns+1]   start of every slot's contributions in ``contrib``
      srow   uint16 [ns]     owned-row index of the slot
      contrib uint16 [nc+1]  shared-memory index of a contribution

    with descriptor entries 6 / 7 / 8 = byte offsets of soff / srow / contrib.
    (Measured alternatives: row-start bit flags and 8-bit counts prefix-summed
    by the warp, 1.25 instead of 4 bytes per slot -- 20-40 % slower, the phase is
    bound by instruction latency, not by the blob's bytes; one 64-bit record per
    slot {CSR position, first contribution, start of the rest} -- fewer loads
    but 8 bytes per slot of shared memory, one CTA less per SM: 0-35 % slower.)

    ``mode="compact"``: slot-parallel, and shared memory holds only the local
    values that some owned slot really reads (a halo element shares one vertex,
    edge or face with the tile's rows, i.e. a fraction of its local matrix --
    in 3-D the halo is several times the tile): per tile element a bit mask
    over the distinct values and the offset of its packed values,

      sbase  int32  [ne_pad]      offset of the element's first stored value
      smask  uint32 [W][ne_pad]   W = ceil(nuniq / 32) mask words

    appended to section A (descriptor entry 15 = byte offset of sbase); value
    ``u`` of element ``el`` lives at ``sbase[el] + popcount(mask bits below u)``."""
    slots_mode = mode in ("slots", "compact")
    compact = mode == "compact"
    # sort_slots (slot-parallel modes): a tile's slots ordered by descending
    # number of contributions, so the 32 slots of a warp step have (nearly) the
    # same trip count -- on 3-D meshes the counts range from 1 to ~24 inside one
    # row and a warp otherwise runs at the pace of its longest slot.  The CSR
    # position then comes per slot (``gpos`` int32 [ns] replaces gdelta + srow).
    sort_slots = sort_slots if slots_mode else False
    dev = indptr.device
    i64 = torch.int64
    N = indptr.numel() - 1
    nnz = cptr.numel() - 1
    nve = vt.shape[0]
    indptr = indptr.to(i64)
    cptr = cptr.to(i64)
    contrib = contrib.to(i64)
    ntiles = (n + E - 1) // E
    ar_t = torch.arange(ntiles, device=dev)
    counts = cptr[1:] - cptr[:-1]
    ar_nnz = torch.arange(nnz, device=dev)
    slot_of_c = torch.repeat_interleave(ar_nnz, counts)
    row_of_slot = torch.repeat_interleave(torch.arange(N, device=dev),
                                          indptr[1:] - indptr[:-1])
    k = contrib % n
    ij = contrib // n
    row_of_c = row_of_slot[slot_of_c]
    if owner_rule == "most":
        # the tile holding most of the row's contributions (ties: the lowest):
        # fewer elements outside the tile have to be recomputed for its rows
        pair, cnt = torch.unique(row_of_c * ntiles + k // E, return_counts=True)
        prow, ptile = pair // ntiles, pair % ntiles
        best = torch.zeros(N, dtype=i64, device=dev)
        best.scatter_reduce_(0, prow, cnt, reduce="amax")
        cand = torch.where(cnt == best[prow], ptile, torch.full_like(ptile, ntiles))
        owner = torch.full((N,), ntiles, dtype=i64, device=dev)
        owner.scatter_reduce_(0, prow, cand, reduce="amin")
    else:
        owner = torch.full((N,), ntiles, dtype=i64, device=dev)
        owner.scatter_reduce_(0, row_of_c, k // E, reduce="amin")
    tile_of_slot = owner[row_of_slot]
    # slots in (tile, row, descending contribution count, CSR position) order:
    # rows of a regular mesh then finish their slots at the same steps, so the
    # lanes of a warp leave the inner accumulation loop together
    csr_pos = ar_nnz - indptr[row_of_slot]           # position of the slot inside its row
    if sort_slots == "bucket":
        # coarse classes (<= 2, 3-4, 5-8, > 8 contributions), CSR order inside a
        # class: most of a warp's stores stay consecutive
        cls = (counts > 2).to(i64) + (counts > 4).to(i64) + (counts > 8).to(i64)
        within = torch.argsort(((3 - cls) << 40) | ar_nnz)
    elif sort_slots:
        within = torch.argsort(((65535 - counts.clamp(max=65535)) << 40) | ar_nnz)
    elif slots_mode or int(csr_pos.max().item()) > 255 or int(counts.max().item()) > 65535:
        within = ar_nnz                               # very long rows: keep CSR order
    else:
        within = torch.argsort((row_of_slot << 32) | ((65535 - counts) << 16) | csr_pos)
    tile_sorted, o2 = torch.sort(tile_of_slot[within], stable=True)
    slot_order = within[o2]
    ns = torch.bincount(tile_of_slot, minlength=ntiles)[:ntiles]
    s0 = torch.cumsum(ns, 0) - ns
    newpos = torch.empty(nnz, dtype=i64, device=dev)
    newpos[slot_order] = ar_nnz
    # contributions in (tile, slot, original) order
    cpos = newpos[slot_of_c]
    cpos_sorted, c_order = torch.sort(cpos, stable=True)
    k_s, ij_s = k[c_order], ij[c_order]
    tile_c = tile_of_slot[slot_of_c][c_order]
    del cpos, slot_of_c, row_of_c
    # tile element lists = unique (tile, element) pairs, sorted by element
    ukey, inv = torch.unique(tile_c * n + k_s, return_inverse=True)
    tile_of_u = ukey // n
    elems = ukey % n
    ne = torch.bincount(tile_of_u, minlength=ntiles)[:ntiles]
    e0 = torch.cumsum(ne, 0) - ne
    eloc = inv - e0[tile_c]
    ne_max = int(ne.max().item())
    ldl = _pad(ne_max, 32) + 1           # odd: consecutive idx rows shift banks
    if nent * ldl > 65535 and not compact:
        raise ValueError("tile too large for 16-bit shared-memory indices")
    if alias is not None:
        # aliased local entries share one shared-memory value (codegen)
        ij_s = torch.as_tensor(alias, dtype=i64, device=dev)[ij_s]
        nent = int(max(alias)) + 1
    max_store, W, smask, sbase = 0, (nent + 31) // 32, None, None
    if compact:
        nte = elems.numel()
        te = e0[tile_c] + eloc                         # tile element of every contribution
        upair, pinv = torch.unique(te * nent + ij_s, return_inverse=True)
        pte, pu = upair // nent, upair % nent          # sorted by (tile element, value)
        cnt = torch.bincount(pte, minlength=nte)
        cum = torch.cumsum(cnt, 0) - cnt               # first pair of every tile element
        tile_pair0 = cum[e0]                           # first pair of every tile
        sbase = cum - tile_pair0[tile_of_u]
        contrib16 = (torch.arange(upair.numel(), device=dev) - tile_pair0[tile_of_u[pte]])[pinv]
        store = torch.zeros(ntiles, dtype=i64, device=dev)
        store.scatter_add_(0, tile_of_u, cnt)
        max_store = int(store.max().item())
        if max_store > 65535:
            raise ValueError("tile too large for 16-bit shared-memory indices")
        smask = torch.zeros(W * nte, dtype=i64, device=dev)
        smask.scatter_add_(0, (pu >> 5) * nte + pte, torch.ones_like(pu) << (pu & 31))
        smask = smask.reshape(W, nte)
    else:
        contrib16 = ij_s * ldl + eloc
    counts_s = counts[slot_order]
    rows_s = row_of_slot[slot_order]
    head = torch.ones(nnz, dtype=i64, device=dev)
    if nnz > 1:
        head[1:] = ((rows_s[1:] != rows_s[:-1]) | (tile_sorted[1:] != tile_sorted[:-1])).to(i64)
    heads_cum = torch.cumsum(head, 0)
    nr = torch.zeros(ntiles, dtype=i64, device=dev)
    nr.scatter_add_(0, tile_sorted, head)
    r0 = torch.cumsum(nr, 0) - nr
    hidx = torch.nonzero(head).squeeze(1)
    gdelta = indptr[rows_s[hidx]] - (hidx - s0[tile_sorted[hidx]])
    # contribution lists: one list per owned row, padded to an even length so
    # that the kernel can read two 16-bit entries with one 32-bit load
    row_c = heads_cum[cpos_sorted] - 1               # owned-row index of every contribution
    n_rows_tot = int(heads_cum[-1].item()) if nnz > 0 else 0
    rc = torch.bincount(row_c, minlength=n_rows_tot)
    first_c = torch.cumsum(rc, 0) - rc
    tile_of_row = tile_sorted[hidx]
    ar_rows = torch.arange(n_rows_tot, device=dev)
    zero_base = nent * ldl                           # 16 zero cells behind the local matrices
    if zero_base + 15 > 32767:
        raise ValueError("tile too large for 15-bit shared-memory indices")
    # last contribution of every slot: the next contribution starts a new slot
    cend = torch.cumsum(counts_s, 0) - 1
    last = torch.zeros(contrib16.numel(), dtype=i64, device=dev)
    last[cend[counts_s > 0]] = 1
    nslots_row = torch.bincount(heads_cum - 1, minlength=n_rows_tot)   # slots per row
    max_slots = int(nslots_row.max().item()) if n_rows_tot else 1
    sched_len = torch.zeros(n_rows_tot, dtype=torch.int32, device=dev)
    out_idx = out_perm = None
    stride = 0
    if schedule and n_rows_tot and max_slots <= 255 and not slots_mode:
        from . import _lib
        import ctypes as C
        lib = _lib.load()
        stride = (2 * int(rc.max().item()) + 8 + 1) // 2 * 2
        ngr = (nr + 15) // 16
        g0 = torch.cumsum(ngr, 0) - ngr
        grp_tile = torch.repeat_interleave(ar_t, ngr)
        grp_i = torch.arange(int(ngr.sum().item()), device=dev) - g0[grp_tile]
        grp_row0 = (r0[grp_tile] + 16 * grp_i).contiguous()
        grp_nrows = torch.clamp(nr[grp_tile] - 16 * grp_i, max=16).to(torch.int32).contiguous()
        idx32 = contrib16.to(torch.int32).contiguous()
        endf = last.to(torch.uint8).contiguous()
        rf = first_c.contiguous()
        rcount = rc.to(torch.int32).contiguous()
        out_idx = torch.zeros(n_rows_tot * stride, dtype=torch.int16, device=dev)
        out_perm = torch.zeros(n_rows_tot * max_slots, dtype=torch.uint8, device=dev)
        on_dev = 1 if dev.type == "cuda" else 0
        stream = None
        if on_dev:
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.spfem_schedule_rows(
            C.c_void_p(idx32.data_ptr()), C.c_void_p(endf.data_ptr()),
            C.c_void_p(rf.data_ptr()), C.c_void_p(rcount.data_ptr()),
            C.c_void_p(grp_row0.data_ptr()), C.c_void_p(grp_nrows.data_ptr()),
            int(grp_row0.numel()), stride, max_slots, zero_base,
            C.c_void_p(out_idx.data_ptr()), C.c_void_p(sched_len.data_ptr()),
            C.c_void_p(out_perm.data_ptr()), 1 if schedule == "inslot" else 0,
            on_dev, stream))
        if on_dev:
            torch.cuda.synchronize(dev)
    sched = sched_len > 0
    rcp = torch.where(sched, sched_len.to(i64), rc + (rc & 1))
    row_k0 = torch.cumsum(rcp, 0) - rcp              # start of every row's (final) list
    nc = torch.zeros(ntiles, dtype=i64, device=dev)
    nc.scatter_add_(0, tile_of_row, rcp)
    c0 = torch.cumsum(nc, 0) - nc
    if int(nc.max().item()) > 65535:
        raise ValueError("more than 65535 contributions in one tile")
    kstart_local = row_k0 - c0[tile_of_row]          # per row, relative to the tile
    # final entry stream (all tiles, rows back to back)
    ent_row = torch.repeat_interleave(ar_rows, rcp)
    ent_t = torch.arange(ent_row.numel(), device=dev) - row_k0[ent_row]
    orig = contrib16 + (last << 15)
    in_orig = ent_t < rc[ent_row]
    src = first_c[ent_row] + torch.where(in_orig, ent_t, torch.zeros_like(ent_t))
    src = torch.clamp(src, max=max(orig.numel() - 1, 0))
    ent_val = torch.where(in_orig, orig[src], torch.full_like(ent_t, zero_base))
    if out_idx is not None:
        sidx = ent_row * stride + torch.clamp(ent_t, max=stride - 1)
        sval = out_idx[sidx].to(i64) & 0xffff
        ent_val = torch.where(sched[ent_row], sval, ent_val)
    ent_tile = tile_of_row[ent_row]
    # slot permutation: CSR position (inside the row) of the j-th finished slot
    slot_row = heads_cum - 1                          # owned-row index of every tile slot
    slot_j = ar_nnz - hidx[slot_row]                  # list position of the slot inside its row
    csr_of_list = csr_pos[slot_order]                 # CSR position of every tile slot
    perm8 = csr_of_list.clone()
    if out_perm is not None:
        pidx = slot_row * max_slots + torch.clamp(slot_j, max=max_slots - 1)
        fin = hidx[slot_row] + out_perm[pidx].to(i64)   # tile slot that finishes j-th
        perm8 = torch.where(sched[slot_row], csr_of_list[torch.clamp(fin, max=nnz - 1)],
                            csr_of_list)
    nchunk = (ns + 31) // 32
    rb0 = torch.cumsum(nchunk, 0) - nchunk
    tot_chunks = int(nchunk.sum().item())
    chunk_tile = torch.repeat_interleave(ar_t, nchunk)
    chunk_idx = torch.arange(tot_chunks, device=dev) - rb0[chunk_tile]
    spos = s0[chunk_tile] + 32 * chunk_idx
    rowbase = heads_cum[spos] - head[spos] - r0[chunk_tile]

    # ---- tile-local vertex numbering ----------------------------------------
    nv_glob = int(vt.max().item()) + 1
    vt_e = vt.to(i64)[:, elems]                      # (nve, total tile elements)
    vkey = (tile_of_u * nv_glob).unsqueeze(0) + vt_e
    uv, vinv = torch.unique(vkey.reshape(-1), return_inverse=True)
    tile_of_v = uv // nv_glob
    vid = uv % nv_glob                               # global id of every tile vertex
    nvt = torch.bincount(tile_of_v, minlength=ntiles)[:ntiles]
    v0 = torch.cumsum(nvt, 0) - nvt
    if int(nvt.max().item()) > 65535:
        raise ValueError("more than 65535 vertices in one tile")
    # number the tile's vertices in order of first use by its elements, so the
    # threads of a warp (consecutive elements) read mostly consecutive
    # coordinates from shared memory (few bank conflicts)
    nte = elems.numel()
    use_pos = (torch.arange(nte, device=dev).unsqueeze(0) * nve
               + torch.arange(nve, device=dev).unsqueeze(1)).reshape(-1)
    first_use = torch.full((uv.numel(),), nte * nve, dtype=i64, device=dev)
    first_use.scatter_reduce_(0, vinv.reshape(-1), use_pos, reduce="amin")
    # first_use increases with the tile index, so one argsort orders every tile
    vorder = torch.argsort(first_use)
    vrank = torch.empty_like(vorder)
    vrank[vorder] = torch.arange(uv.numel(), device=dev)
    tile_of_v = tile_of_v[vorder]
    vid = vid[vorder]
    vloc = vrank[vinv].reshape(nve, -1) - v0[tile_of_u].unsqueeze(0)
    # ---- pack the per-tile blobs ------------------------------------------
    ne_pad = (ne + 7) // 8 * 8
    nvt_pad = (nvt + 1) // 2 * 2
    off_el = nve * ne_pad * 2
    off_pc = off_el + (ne_pad * 4 if with_el else 0)
    bytes_a = off_pc + dim * nvt_pad * 8
    off_base = torch.zeros_like(ne)
    if compact:
        off_base = bytes_a
        bytes_a = (off_base + (1 + W) * 4 * ne_pad + 15) // 16 * 16
    if slots_mode:
        nc = torch.bincount(tile_c, minlength=ntiles)[:ntiles]
        c0 = torch.cumsum(nc, 0) - nc
        if int(nc.max().item()) > 65535 or (not sort_slots and int(nr.max().item()) > 65535):
            raise ValueError("more than 65535 contributions / rows in one tile")
        if sort_slots:
            off_cptr = (ns * 4 + 15) // 16 * 16              # soff (behind gpos)
            off_rowbase = torch.zeros_like(nr)
            off_contrib = off_cptr + ((ns + 1) * 2 + 15) // 16 * 16
        else:
            off_cptr = (nr * 4 + 15) // 16 * 16                  # soff
            off_rowbase = off_cptr + ((ns + 1) * 2 + 15) // 16 * 16   # srow
            off_contrib = off_rowbase + (ns * 2 + 15) // 16 * 16
        bytes_b = off_contrib + ((nc + 1) * 2 + 15) // 16 * 16
    else:
        off_rowbase = torch.zeros_like(nr)
        off_cptr = ((nr + 1) * 8 + 15) // 16 * 16          # = offset of the perm section
        off_contrib = off_cptr + (ns + 15) // 16 * 16
        bytes_b = off_contrib + (nc * 2 + 15) // 16 * 16
    tile_bytes = bytes_a + bytes_b
    blob_off = torch.cumsum(tile_bytes, 0) - tile_bytes
    total = int(tile_bytes.sum().item())
    if int((blob_off // 16).max().item()) > 2 ** 31 - 1:
        raise ValueError("tile blob exceeds 32 GB")
    blob = torch.zeros(total, dtype=torch.uint8, device=dev)
    b32 = blob.view(torch.int32)
    b16 = blob.view(torch.int16)
    # section A: local vertex table, optional element ids
    el_local = torch.arange(elems.numel(), device=dev) - e0[tile_of_u]
    base16 = (blob_off // 2)[tile_of_u]
    npad_u = ne_pad[tile_of_u]
    for r in range(nve):
        b16[base16 + r * npad_u + el_local] = vloc[r].to(torch.int16)
    if with_el:
        b32[((blob_off + off_el) // 4)[tile_of_u] + el_local] = elems.to(torch.int32)
    if compact:
        bb = ((blob_off + off_base) // 4)[tile_of_u]
        b32[bb + el_local] = sbase.to(torch.int32)
        for w in range(W):
            b32[bb + (1 + w) * npad_u + el_local] = _wrap32(torch, smask[w])
    # destination (in float64 units) of coordinate d of every tile vertex:
    # pc_index + d * pc_stride
    v_local = torch.arange(vid.numel(), device=dev) - v0[tile_of_v]
    if dim == 2:
        # interleaved (x, y) pairs: one 128-bit shared-memory load per vertex
        pc_index = ((blob_off + off_pc) // 8)[tile_of_v] + 2 * v_local
        pc_stride = torch.ones_like(v_local)
    else:
        pc_index = ((blob_off + off_pc) // 8)[tile_of_v] + v_local
        pc_stride = nvt_pad[tile_of_v]
    # section B
    bbase = blob_off + bytes_a
    hrow_local = heads_cum[hidx] - 1 - r0[tile_sorted[hidx]]
    slot_local = ar_nnz - s0[tile_sorted]
    if slots_mode:
        if sort_slots:
            b32[(bbase // 4)[tile_sorted] + slot_local] = slot_order.to(torch.int32)
        else:
            b32[(bbase // 4)[tile_sorted[hidx]] + hrow_local] = gdelta.to(torch.int32)
        cstart = torch.cumsum(counts_s, 0) - counts_s        # first contribution of every slot
        b16[((bbase + off_cptr) // 2)[tile_sorted] + slot_local] = \
            _wrap16(torch, cstart - c0[tile_sorted])
        b16[((bbase + off_cptr) // 2) + ns] = _wrap16(torch, nc)
        if not sort_slots:
            b16[((bbase + off_rowbase) // 2)[tile_sorted] + slot_local] = \
                _wrap16(torch, heads_cum - 1 - r0[tile_sorted])
        cb_base = ((bbase + off_contrib) // 2)
        ar_c = torch.arange(contrib16.numel(), device=dev)
        b16[cb_base[tile_c] + (ar_c - c0[tile_c])] = contrib16.to(torch.int16)
    else:
        ri_base = (bbase // 4)[tile_sorted[hidx]] + 2 * hrow_local
        b32[ri_base] = gdelta.to(torch.int32)
        packed = slot_local[hidx] + (kstart_local << 16)
        b32[ri_base + 1] = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).to(torch.int32)
        # terminating entry of every tile: {0, ns | nc << 16}
        endp = ns + (nc << 16)
        b32[(bbase // 4) + 2 * nr + 1] = torch.where(endp >= 2 ** 31, endp - 2 ** 32, endp).to(torch.int32)
        cb_base = ((bbase + off_contrib) // 2)
        b16[cb_base[ent_tile] + (row_k0[ent_row] + ent_t - c0[ent_tile])] = ent_val.to(torch.int16)
        blob[(bbase + off_cptr)[tile_sorted] + slot_local] = perm8.to(torch.uint8)
    z = torch.zeros_like(ne)
    desc = torch.stack([blob_off // 16, bytes_a, bytes_b, ne, ne_pad, ns, off_cptr,
                        off_rowbase, off_contrib, nr, nc, nvt_pad, off_pc,
                        off_el if with_el else z - 1, nvt, off_base], 1)
    if int(desc.max().item()) > 2 ** 31 - 1:
        raise ValueError("tile plan offsets exceed 32 bits")
    return {
        "ntiles": ntiles, "ldl": ldl, "E": E, "nve": nve, "nuniq": nent, "mode": mode,
        "max_store": max_store, "W": W, "sort_slots": sort_slots,
        "desc": desc.to(torch.int32).contiguous(),
        "blob": blob,
        "max_bytes_a": int(bytes_a.max().item()),
        "max_bytes_b": int(bytes_b.max().item()),
        "max_ns": int(ns.max().item()), "max_nr": int(nr.max().item()),
        "max_ne": ne_max,
        "tile_elems": (elems, e0, ne),
        "vid": vid, "pc_index": pc_index, "pc_stride": pc_stride, "dim": dim,
        "tile_vertices": float(nvt.sum().item()),
        "scheduled_rows": float(sched.sum().item()) / max(n_rows_tot, 1),
        "list_growth": float(rcp.sum().item()) / max(float(rc.sum().item()), 1.0),
        "halo_factor": float(ne.sum().item()) / float(n),
        "blob_bytes": total,
    }


def emulate(plan, local, nnz):
    """NumPy emulation of phase B of the tile kernel (tests only): ``local`` is
    the entry-major ``(nent, n)`` array of local matrices.  Reads only the
    packed blob, exactly like the kernel."""
    desc = plan["desc"].cpu().numpy().astype(np.int64)
    blob = plan["blob"].cpu().numpy()
    ldl, nve = plan["ldl"], plan["nve"]
    nent = local.shape[0]
    values = np.full(nnz, np.nan)
    written = np.zeros(nnz, dtype=np.int64)
    all_elems = plan["tile_elems"][0].cpu().numpy()
    e0s = plan["tile_elems"][1].cpu().numpy()
    for ti, (off16, ba, bb, ne, ne_pad, ns, o_cp, o_rb, o_cb, nr, nc, *_r) in enumerate(desc):
        B = blob[off16 * 16 + ba: off16 * 16 + ba + bb]
        elems = all_elems[e0s[ti]: e0s[ti] + ne]
        if plan.get("mode") == "compact":
            A = blob[off16 * 16: off16 * 16 + ba]
            ob, W = _r[4], plan["W"]
            sb = A[ob: ob + 4 * ne_pad].view(np.int32).astype(np.int64)
            sm = A[ob + 4 * ne_pad: ob + 4 * (1 + W) * ne_pad].view(np.uint32).reshape(W, ne_pad)
            sLz = np.full(plan["max_store"], np.nan)
            for el in range(ne):
                pos = sb[el]
                for u in range(nent):
                    if (int(sm[u >> 5, el]) >> (u & 31)) & 1:
                        sLz[pos] = local[u, elems[el]]
                        pos += 1
        else:
            sLz = np.zeros(nent * ldl + 16)
            for idx in range(nent):
                sLz[idx * ldl: idx * ldl + ne] = local[idx, elems]
        if plan.get("mode") in ("slots", "compact"):
            srt = plan.get("sort_slots", False)
            gd = B[:(ns if srt else nr) * 4].view(np.int32).astype(np.int64)
            soff = B[o_cp:o_cp + 2 * (ns + 1)].view(np.uint16).astype(np.int64)
            srow = B[o_rb:o_rb + 2 * ns].view(np.uint16).astype(np.int64)
            contrib = B[o_cb:o_cb + 2 * nc].view(np.uint16).astype(np.int64)
            for sl in range(ns):
                acc = 0.0
                for kk in range(soff[sl], soff[sl + 1]):
                    acc += sLz[contrib[kk]]
                g = gd[sl] if srt else gd[srow[sl]] + sl
                values[g] = acc
                written[g] += 1
            continue
        rinfo = B[:(nr + 1) * 8].view(np.int32).astype(np.int64).reshape(nr + 1, 2)
        perm = B[o_cp:o_cp + ns].astype(np.int64)
        contrib = B[o_cb:o_cb + 2 * nc].view(np.uint16).astype(np.int64)
        for row in range(nr):
            packed = rinfo[row, 1] & 0xffffffff
            kend = (rinfo[row + 1, 1] & 0xffffffff) >> 16
            s = packed & 0xffff
            g0 = rinfo[row, 0] + s
            acc = 0.0
            for kk in range(packed >> 16, kend):
                c = contrib[kk]
                acc += sLz[c & 0x7fff]
                if c & 0x8000:
                    g = g0 + perm[s]
                    values[g] = acc
                    written[g] += 1
                    acc = 0.0
                    s += 1
    return values, written


def fill_coordinates(plan, p):
    """Copy the vertex coordinates ``p`` (dim, nv) into the tiles' private
    coordinate sections (device-side gather; call after every change of p)."""
    b64 = plan["blob"].view(p.dtype)
    for d in range(plan["dim"]):
        b64[plan["pc_index"] + d * plan["pc_stride"]] = p[d][plan["vid"]]


def check_vertex_table(plan, vt, p):
    """Tests only: section A of every blob holds, through the tile-local vertex
    ids, the coordinates of the vertices of the tile's elements."""
    desc = plan["desc"].cpu().numpy().astype(np.int64)
    blob = plan["blob"].cpu().numpy()
    nve, dim = plan["nve"], plan["dim"]
    vt = np.asarray(vt)
    all_elems = plan["tile_elems"][0].cpu().numpy()
    e0s = plan["tile_elems"][1].cpu().numpy()
    for ti, d in enumerate(desc):
        off16, ba, ne, ne_pad, nvt_pad, off_pc = d[0], d[1], d[3], d[4], d[11], d[12]
        A = blob[off16 * 16: off16 * 16 + ba]
        loc = A[:nve * ne_pad * 2].view(np.uint16).reshape(nve, ne_pad)[:, :ne]
        pc = A[off_pc: off_pc + dim * nvt_pad * 8].view(np.float64)
        pc = pc.reshape(nvt_pad, 2).T if dim == 2 else pc.reshape(dim, nvt_pad)
        elems = all_elems[e0s[ti]: e0s[ti] + ne]
        for r in range(nve):
            for k in range(dim):
                if not np.array_equal(pc[k, loc[r]], p[k, vt[r, elems]]):
                    return False
    return True


def bank_conflicts(plan):
    """Tests only: average number of shared-memory wavefronts per half-warp
    step of phase B1 (1.0 = conflict free) and the schedule's length growth."""
    desc = plan["desc"].cpu().numpy().astype(np.int64)
    blob = plan["blob"].cpu().numpy()
    waves = steps = 0
    for d in desc:
        off16, ba, bb, ns, o_cb, nr, nc = d[0], d[1], d[2], d[5], d[8], d[9], d[10]
        B = blob[off16 * 16 + ba: off16 * 16 + ba + bb]
        rinfo = B[:(nr + 1) * 8].view(np.int32).astype(np.int64).reshape(nr + 1, 2)
        contrib = B[o_cb:o_cb + 2 * nc].view(np.uint16).astype(np.int64)
        k0 = (rinfo[:, 1] & 0xffffffff) >> 16
        for r0 in range(0, nr, 16):
            rows = range(r0, min(nr, r0 + 16))
            L = max(k0[r + 1] - k0[r] for r in rows)
            for t in range(L):
                addr = set(int(contrib[k0[r] + t] & 0x7fff) for r in rows
                           if t < k0[r + 1] - k0[r])
                banks = np.bincount([a & 15 for a in addr], minlength=16)
                waves += banks.max()
                steps += 1
    return waves / max(steps, 1)
