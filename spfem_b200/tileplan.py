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

Per tile (descriptor ``e0, ne, cp0, ns, c0, r0, rb0``):
  elems[e0 : e0+ne]        element ids whose local matrices the CTA computes
  cptr[cp0 : cp0+ns+1]     uint16 start of each owned slot's contribution list
                           (relative to c0), bit 15 marks the first slot of a row
  contrib[c0 : ...]        uint16 shared-memory index ``idx*ldl + el``
  rowbase[rb0 + s//32]     number of owned rows that start before slot 32*(s//32)
  gdelta[r0 + row]         CSR position of the row's first slot minus its tile
                           slot index
"""
import numpy as np


def build_tile_plan(torch, indptr, cptr, contrib, n, nent, E):
    """All arguments are tensors on one device; returns a dict of tensors."""
    dev = indptr.device
    i64 = torch.int64
    N = indptr.numel() - 1
    nnz = cptr.numel() - 1
    indptr = indptr.to(i64)
    cptr = cptr.to(i64)
    contrib = contrib.to(i64)
    ntiles = (n + E - 1) // E
    counts = cptr[1:] - cptr[:-1]
    ar_nnz = torch.arange(nnz, device=dev)
    slot_of_c = torch.repeat_interleave(ar_nnz, counts)
    row_of_slot = torch.repeat_interleave(torch.arange(N, device=dev),
                                          indptr[1:] - indptr[:-1])
    k = contrib % n
    ij = contrib // n
    row_of_c = row_of_slot[slot_of_c]
    owner = torch.full((N,), ntiles, dtype=i64, device=dev)
    owner.scatter_reduce_(0, row_of_c, k // E, reduce="amin")
    tile_of_slot = owner[row_of_slot]
    # slots in (tile, CSR position) order
    tile_sorted, slot_order = torch.sort(tile_of_slot, stable=True)
    ns = torch.bincount(tile_of_slot, minlength=ntiles)[:ntiles]
    s0 = torch.cumsum(ns, 0) - ns
    newpos = torch.empty(nnz, dtype=i64, device=dev)
    newpos[slot_order] = ar_nnz
    # contributions in (tile, slot, original) order
    cpos = newpos[slot_of_c]
    _, c_order = torch.sort(cpos, stable=True)
    k_s, ij_s = k[c_order], ij[c_order]
    tile_c = tile_of_slot[slot_of_c][c_order]
    # tile element lists = unique (tile, element) pairs, sorted by element
    ukey, inv = torch.unique(tile_c * n + k_s, return_inverse=True)
    tile_of_u = ukey // n
    elems = (ukey % n).to(torch.int32)
    ne = torch.bincount(tile_of_u, minlength=ntiles)[:ntiles]
    e0 = torch.cumsum(ne, 0) - ne
    eloc = inv - e0[tile_c]
    ldl = int(ne.max().item())
    ldl = (ldl + 31) // 32 * 32 + 1          # odd multiple: spreads banks across idx
    if nent * ldl > 65535:
        raise ValueError("tile too large for 16-bit shared-memory indices")
    contrib16 = (ij_s * ldl + eloc).to(torch.int32)
    # per-slot contribution list starts (relative to the tile's base)
    counts_s = counts[slot_order]
    cstart = torch.cumsum(counts_s, 0) - counts_s
    nc = torch.zeros(ntiles, dtype=i64, device=dev)
    nc.scatter_add_(0, tile_sorted, counts_s)
    c0 = torch.cumsum(nc, 0) - nc
    lstart = cstart - c0[tile_sorted]
    if int(nc.max().item()) > 32767:
        raise ValueError("more than 32767 contributions in one tile")
    rows_s = row_of_slot[slot_order]
    head = torch.ones(nnz, dtype=i64, device=dev)
    if nnz > 1:
        head[1:] = ((rows_s[1:] != rows_s[:-1]) | (tile_sorted[1:] != tile_sorted[:-1])).to(i64)
    cptr16 = torch.empty(nnz + ntiles, dtype=torch.int32, device=dev)
    cptr16[ar_nnz + tile_sorted] = (lstart + (head << 15)).to(torch.int32)
    cp0 = s0 + torch.arange(ntiles, device=dev)
    cptr16[cp0 + ns] = nc.to(torch.int32)
    # rows
    heads_cum = torch.cumsum(head, 0)
    nr = torch.zeros(ntiles, dtype=i64, device=dev)
    nr.scatter_add_(0, tile_sorted, head)
    r0 = torch.cumsum(nr, 0) - nr
    hidx = torch.nonzero(head).squeeze(1)
    gdelta = (slot_order[hidx] - (hidx - s0[tile_sorted[hidx]])).to(torch.int32)
    # rowbase: heads strictly before every 32nd slot of each tile
    nchunk = (ns + 31) // 32
    rb0 = torch.cumsum(nchunk, 0) - nchunk
    tot_chunks = int(nchunk.sum().item())
    chunk_tile = torch.repeat_interleave(torch.arange(ntiles, device=dev), nchunk)
    chunk_idx = torch.arange(tot_chunks, device=dev) - rb0[chunk_tile]
    spos = s0[chunk_tile] + 32 * chunk_idx
    before = heads_cum[spos] - head[spos] - r0[chunk_tile]
    rowbase = before.to(torch.int32)
    desc = torch.stack([e0, ne, cp0, ns, c0, r0, rb0, torch.zeros_like(e0)], 1)
    if int(desc.max().item()) > 2 ** 31 - 1:
        raise ValueError("tile plan offsets exceed 32 bits")
    return {
        "ntiles": ntiles, "ldl": ldl, "E": E,
        "desc": desc.to(torch.int32).contiguous(),
        "elems": elems.contiguous(),
        "cptr": cptr16.to(torch.int16).contiguous(),
        "contrib": contrib16.to(torch.int16).contiguous(),
        "rowbase": rowbase.to(torch.int16).contiguous(),
        "gdelta": gdelta.contiguous(),
        "halo_factor": float(ne.sum().item()) / float(n),
        "bytes": None,
    }


def emulate(plan, local, n, nnz):
    """NumPy emulation of phase B of the tile kernel (tests only): ``local`` is
    the entry-major ``(nent, n)`` array of local matrices."""
    desc = plan["desc"].cpu().numpy()
    elems = plan["elems"].cpu().numpy()
    cptr = plan["cptr"].cpu().numpy().view(np.uint16).astype(np.int64)
    contrib = plan["contrib"].cpu().numpy().view(np.uint16).astype(np.int64)
    rowbase = plan["rowbase"].cpu().numpy().view(np.uint16).astype(np.int64)
    gdelta = plan["gdelta"].cpu().numpy().astype(np.int64)
    ldl = plan["ldl"]
    nent = local.shape[0]
    values = np.full(nnz, np.nan)
    written = np.zeros(nnz, dtype=np.int64)
    for (e0, ne, cp0, ns, c0, r0, rb0, _) in desc:
        sL = np.zeros(nent * ldl)
        for idx in range(nent):
            sL[idx * ldl: idx * ldl + ne] = local[idx, elems[e0:e0 + ne]]
        cp = cptr[cp0:cp0 + ns + 1]
        heads = (cp[:ns] >> 15) & 1
        starts = np.concatenate((cp[:ns] & 0x7fff, [cp[ns] & 0x7fff]))
        for s in range(ns):
            chunk = s >> 5
            row = rowbase[rb0 + chunk] + int(heads[32 * chunk: s + 1].sum()) - 1
            acc = 0.0
            for kk in range(starts[s], starts[s + 1]):
                acc += sL[contrib[c0 + kk]]
            g = gdelta[r0 + row] + s
            values[g] = acc
            written[g] += 1
    return values, written
