# -*- coding: utf-8 -*-
"""Device-side construction of the vertex-fan plan (``codegen.FAN_KERNEL``).

Written with PyTorch index operations so that it runs on the GPU next to the
pattern build: set-up plumbing, not the hot path.  ``tests/fanplan_ref.py`` is
the readable NumPy statement of the same plan (byte-identical blobs; the
yardstick of ``tests/test_fanplan.py``)."""

MAXN = 12          # ring neighbours per vertex the general kernel supports


class FanPlanError(ValueError):
    """The mesh / pattern is not vertex-fan shaped (the caller falls back to the
    element-tile kernel)."""


def fill_fan_coordinates(plan, p):
    """Copy the vertex coordinates ``p`` (2, nv) into the tiles' private
    interleaved (x, y) sections (call after every change of ``p``)."""
    b64 = plan["blob"].view(p.dtype)
    for d in range(2):
        b64[plan["pc_index"] + d] = p[d][plan["vid"]]


def _lexsort(torch, keys):
    """Indices sorting by the last key first (``np.lexsort`` convention)."""
    order = torch.argsort(keys[0], stable=True)
    for k in keys[1:]:
        order = order[torch.argsort(k[order], stable=True)]
    return order


def _morton_vertices(torch, p, bits=20):
    lo = p.min(dim=1, keepdim=True).values
    ext = p.max(dim=1, keepdim=True).values - lo
    ext = torch.where(ext == 0, torch.ones_like(ext), ext)
    q = torch.clamp(((p - lo) / ext * (2 ** bits - 1)).to(torch.int64), max=2 ** bits - 1)
    key = torch.zeros(p.shape[1], dtype=torch.int64, device=p.device)
    for b in range(bits):
        for d in range(p.shape[0]):
            key |= ((q[d] >> b) & 1) << (b * p.shape[0] + d)
    return torch.argsort(key, stable=True)


def vertex_rings(torch, p, t, t2f, f2t):
    """Torch version of :func:`fanplan.vertex_rings` (same outputs)."""
    dev = p.device
    i64 = torch.int64
    nt, nv = t.shape[1], p.shape[1]
    ar = torch.arange(nt, device=dev)
    x, y = p[0][t], p[1][t]
    det = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0])
    if bool((det == 0).any()):
        raise FanPlanError("degenerate triangle")
    ccw = det > 0
    li = torch.arange(3, device=dev)[:, None]
    nxt = torch.where(ccw[None, :], (li + 1) % 3, (li + 2) % 3)
    prv = torch.where(ccw[None, :], (li + 2) % 3, (li + 1) % 3)
    LF = torch.tensor([[-1, 0, 2], [0, -1, 1], [2, 1, -1]], device=dev)

    def across(i, j):
        fid = t2f[LF[i, j], ar]
        f0, f1 = f2t[0, fid], f2t[1, fid]
        return torch.where(f0 == ar, f1, f0)

    next_corner = torch.full((3 * nt,), -1, dtype=i64, device=dev)
    has_prev = torch.zeros(3 * nt, dtype=torch.bool, device=dev)
    for i in range(3):
        ii = torch.full((nt,), i, dtype=i64, device=dev)
        r = t[i]
        tn = across(ii, prv[i])
        ok = tn >= 0
        tnc = torch.where(ok, tn, torch.zeros_like(tn))
        iloc = torch.argmax((t[:, tnc] == r[None, :]).to(torch.int8), dim=0)
        next_corner[i * nt + ar] = torch.where(ok, iloc * nt + tnc, torch.full_like(tn, -1))
        has_prev[i * nt + ar] = across(ii, nxt[i]) >= 0
    corner_vertex = t.reshape(-1)
    valence = torch.bincount(corner_vertex, minlength=nv)
    if int(valence.min()) == 0:
        raise FanPlanError("vertex without elements")
    K = int(valence.max())
    if K > MAXN:
        raise FanPlanError("vertex valence %d > %d" % (K, MAXN))
    cnxt = nxt.reshape(-1)
    cprv = prv.reshape(-1)
    corner_T = ar.repeat(3)
    a_of = t[cnxt, corner_T]
    b_of = t[cprv, corner_T]
    ang = torch.atan2(p[1][a_of] - p[1][corner_vertex], p[0][a_of] - p[0][corner_vertex])
    ordk = _lexsort(torch, [ang, corner_vertex])
    start = ordk[torch.searchsorted(corner_vertex[ordk].contiguous(),
                                    torch.arange(nv, device=dev))]
    nopred = torch.nonzero(~has_prev).squeeze(1)
    start[corner_vertex[nopred]] = nopred
    nb = torch.full((K + 1, nv), -1, dtype=i64, device=dev)
    nfan = torch.zeros(nv, dtype=i64, device=dev)
    closed = torch.zeros(nv, dtype=torch.bool, device=dev)
    cur = start.clone()
    alive = torch.ones(nv, dtype=torch.bool, device=dev)
    last_b = torch.full((nv,), -1, dtype=i64, device=dev)
    for k in range(K):
        idx = torch.nonzero(alive).squeeze(1)
        c = cur[idx]
        nb[k, idx] = a_of[c]
        last_b[idx] = b_of[c]
        nfan[idx] += 1
        nc = next_corner[c]
        back = nc == start[idx]
        closed[idx[back]] = True
        stop = back | (nc < 0)
        alive[idx[stop]] = False
        cur[idx] = nc
    if bool(alive.any()) or not torch.equal(nfan, valence):
        raise FanPlanError("non-manifold vertex fan")
    opn = torch.nonzero(~closed).squeeze(1)
    nb[nfan[opn], opn] = last_b[opn]
    return nb, nfan, closed


def _wrap32(torch, x):
    return torch.where(x >= 2 ** 31, x - 2 ** 32, x).to(torch.int32)


def build_fan_plan(torch, p, t, t2f, f2t, indptr, indices, rows_per_tile=128, tiny=True,
                   row_order="id"):
    """Torch version of :func:`fanplan.build_fan_plan`; all tensor arguments on
    one device (int64 topology, float64 coordinates).  Returns the same dict with
    tensors instead of arrays."""
    dev = p.device
    i64 = torch.int64
    nv = p.shape[1]
    if p.shape[0] != 2 or indptr.numel() != nv + 1:
        raise FanPlanError("not a scalar vertex-based pattern on a 2-D mesh")
    nb, nfan, closed = vertex_rings(torch, p, t, t2f, f2t)
    nn = nfan + (~closed).to(i64)
    if int(nn.max()) > MAXN:
        raise FanPlanError("more than %d ring neighbours" % MAXN)
    indptr = indptr.to(i64)
    indices = indices.to(i64)
    rowlen = indptr[1:] - indptr[:-1]
    if not torch.equal(rowlen, nn + 1):
        raise FanPlanError("pattern rows are not vertex stars")
    if int(indptr[-1]) > 2 ** 31 - 1:
        raise FanPlanError("nnz exceeds 32-bit positions")
    arv = torch.arange(nv, device=dev)
    rowkey = torch.repeat_interleave(arv, rowlen) * nv + indices
    K1 = nb.shape[0]
    vv = arv.unsqueeze(0).expand(K1, nv)
    valid = nb >= 0
    gpos = torch.searchsorted(rowkey, vv[valid] * nv + nb[valid])
    if not torch.equal(indices[torch.clamp(gpos, max=indices.numel() - 1)], nb[valid]):
        raise FanPlanError("ring neighbour missing from the pattern")
    pos = torch.zeros((K1, nv), dtype=i64, device=dev)
    pos[valid] = gpos - indptr[vv[valid]]
    dpos = torch.searchsorted(rowkey, arv * nv + arv) - indptr[:-1]
    nb12 = torch.full((MAXN, nv), -1, dtype=i64, device=dev)
    pos12 = torch.zeros((MAXN, nv), dtype=i64, device=dev)
    kk = min(K1, MAXN)
    nb12[:kk], pos12[:kk] = nb[:kk], pos[:kk]
    compact = int(nn.max()) <= 8
    W = 8 if compact else MAXN
    R = int(rows_per_tile)
    ntiles = (nv + R - 1) // R
    vperm = _morton_vertices(torch, p)
    tile_of_row = arv // R
    if row_order == "id":
        # rows of a tile in CSR order: rows with consecutive ids are neighbours in a
        # warp, so the sectors two of them share are written by one store request
        order = _lexsort(torch, [vperm, tile_of_row])
    else:
        order = _lexsort(torch, [p[0][vperm], p[1][vperm], tile_of_row])
    rows = vperm[order]
    nr = torch.bincount(tile_of_row, minlength=ntiles)
    r0 = torch.cumsum(nr, 0) - nr
    row_local = arv - r0[tile_of_row]
    tile_of_vertex = torch.empty(nv, dtype=i64, device=dev)
    tile_of_vertex[rows] = tile_of_row
    local_of_vertex = torch.empty(nv, dtype=i64, device=dev)
    local_of_vertex[rows] = row_local
    nbr = nb12[:W][:, rows]
    ok = nbr >= 0
    nbr0 = torch.where(ok, nbr, torch.zeros_like(nbr))
    tl = tile_of_row.unsqueeze(0).expand(W, nv)
    inside = ok & (tile_of_vertex[nbr0] == tl)
    out = ok & ~inside
    hk = tl[out] * nv + nbr[out]
    uh = torch.unique(hk)
    h_tile = uh // nv
    nh = torch.bincount(h_tile, minlength=ntiles)
    h0 = torch.cumsum(nh, 0) - nh
    nvt = nr + nh
    if int(nvt.max()) > 65535:
        raise FanPlanError("tile with more than 65535 vertices")
    loc = torch.zeros((W, nv), dtype=i64, device=dev)
    loc[inside] = local_of_vertex[nbr[inside]]
    hv = uh % nv
    hord = _lexsort(torch, [p[0][hv], p[1][hv], h_tile])
    hrank = torch.empty(uh.numel(), dtype=i64, device=dev)
    hrank[hord] = torch.arange(uh.numel(), device=dev)
    hidx = torch.searchsorted(uh, hk)
    loc[out] = nr[tl[out]] + (hrank[hidx] - h0[tl[out]])
    lens = rowlen[rows]
    ns = torch.zeros(ntiles, dtype=i64, device=dev)
    ns.scatter_add_(0, tile_of_row, lens)
    cl = torch.cumsum(lens, 0) - lens
    s0 = cl - cl[r0[tile_of_row]]
    w0 = indptr[rows]
    w1 = s0 | (nfan[rows] << 16) | (closed[rows].to(i64) << 20) | (dpos[rows] << 24)
    nbl = loc.t()                                           # (nv, W)
    psl = torch.where(ok, pos12[:W][:, rows], torch.zeros_like(nbr)).t()

    def pair(a, b):
        return _wrap32(torch, a | (b << 16))

    def quad(a, b, c, d):
        return _wrap32(torch, a | (b << 8) | (c << 16) | (d << 24))
    z = torch.zeros(nv, dtype=torch.int32, device=dev)
    tiny = bool(tiny) and compact and int(nvt.max()) <= 256
    nib = None
    if tiny:
        plane0 = torch.stack([_wrap32(torch, w0), _wrap32(torch, w1),
                              quad(nbl[:, 0], nbl[:, 1], nbl[:, 2], nbl[:, 3]),
                              quad(nbl[:, 4], nbl[:, 5], nbl[:, 6], nbl[:, 7])], 1)
        nib = torch.zeros(nv, dtype=i64, device=dev)
        for q in range(8):
            nib |= psl[:, q] << (4 * q)
    else:
        plane0 = torch.stack([_wrap32(torch, w0), _wrap32(torch, w1),
                              pair(nbl[:, 0], nbl[:, 1]), pair(nbl[:, 2], nbl[:, 3])], 1)
    planes = [plane0]
    if tiny:
        pass
    elif compact:
        planes.append(torch.stack([pair(nbl[:, 4], nbl[:, 5]), pair(nbl[:, 6], nbl[:, 7]),
                                   quad(psl[:, 0], psl[:, 1], psl[:, 2], psl[:, 3]),
                                   quad(psl[:, 4], psl[:, 5], psl[:, 6], psl[:, 7])], 1))
    else:
        planes.append(torch.stack([pair(nbl[:, 4], nbl[:, 5]), pair(nbl[:, 6], nbl[:, 7]),
                                   pair(nbl[:, 8], nbl[:, 9]), pair(nbl[:, 10], nbl[:, 11])], 1))
        planes.append(torch.stack([quad(psl[:, 0], psl[:, 1], psl[:, 2], psl[:, 3]),
                                   quad(psl[:, 4], psl[:, 5], psl[:, 6], psl[:, 7]),
                                   quad(psl[:, 8], psl[:, 9], psl[:, 10], psl[:, 11]), z], 1))
    bytes_a = (nvt * 16 + 15) // 16 * 16
    bytes_b = nr * 16 * len(planes)
    if tiny:
        bytes_b = bytes_b + (nr * 4 + 15) // 16 * 16
    tb = bytes_a + bytes_b
    off = torch.cumsum(tb, 0) - tb
    total = int(tb.sum())
    blob32 = torch.zeros(total // 4, dtype=torch.int32, device=dev)
    a4 = torch.arange(4, device=dev)
    for k, pl in enumerate(planes):
        dk = ((off + bytes_a + k * nr * 16)[tile_of_row] + 16 * row_local) // 4
        blob32[dk[:, None] + a4[None, :]] = pl
    if tiny:
        blob32[((off + bytes_a + nr * 16)[tile_of_row] + 4 * row_local) // 4] = _wrap32(torch, nib)
    zt = torch.zeros_like(nr)
    desc = torch.stack([off // 16, bytes_a, bytes_b, nr, ns, nvt, zt, zt], 1)
    if int(desc.max()) > 2 ** 31 - 1:
        raise FanPlanError("fan plan offsets exceed 32 bits")
    vid = torch.cat((rows, hv))
    vtile = torch.cat((tile_of_row, h_tile))
    vloc = torch.cat((row_local, nr[h_tile] + hrank - h0[h_tile]))
    pc_index = off[vtile] // 8 + 2 * vloc
    return {
        "ntiles": ntiles, "rows_per_tile": R,
        "desc": desc.to(torch.int32).contiguous(), "blob": blob32.view(torch.uint8),
        "vid": vid, "pc_index": pc_index,
        "max_a": int(tb.max()), "max_b": 0, "max_ns": int(ns.max()),
        "halo_factor": float(nvt.sum()) / nv,
        "max_neighbours": int(nn.max()),
        "format": "tiny" if tiny else ("fan8" if compact else "wide"),
    }
