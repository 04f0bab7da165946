# -*- coding: utf-8 -*-
"""Plan of the vertex-fan kernel (``codegen.FAN_KERNEL``).

For scalar P1 on triangles (one DOF per vertex) with element-constant
coefficients every CSR row is a mesh vertex ``r`` and its columns are ``r`` and
its ring neighbours.  The kernel gives each row to one thread that walks the
triangles around ``r`` in counter-clockwise order and recomputes the local row
of ``r`` from vertex coordinates -- no local matrices in shared memory.  This
module builds, once per mesh (host NumPy; set-up, not the hot path):

* the ring order of the triangles / neighbours around every vertex, from the
  mesh's own topology (``t2f`` / ``f2t``),
* tiles of ``R`` Morton-consecutive rows with their tile-local vertex numbering
  (the rows themselves first, then the halo vertices),
* one 16-byte aligned blob per tile (brought into shared memory by one TMA
  bulk copy): section A = interleaved ``(x, y)`` coordinates of the tile's
  vertices; section B = the row records as planes (structure of arrays, so the
  threads of a warp read consecutive 16 byte words): plane 0 =
  ``{g0, s0 | nfan << 16 | closed << 20 | posdiag << 24, nbr0..3}``, plane 1 =
  ``{nbr4..7, pos0..7}`` (<= 8 neighbours) or ``{nbr4..11}`` + plane 2 =
  ``{pos0..11}``: ring neighbours as 16-bit tile-local vertex ids and the CSR
  position (inside the row) of every neighbour as bytes.

Descriptor per tile (int32 x 8): blob offset / 16, bytes A, bytes B, rows,
slots, vertices, 0, 0.
"""
import numpy as np

ROW_DTYPE = np.dtype([("g0", "<i4"), ("s0", "<u2"), ("nfan", "u1"), ("closed", "u1"),
                      ("nbr", "<u2", (12,)), ("pos", "u1", (12,)),
                      ("posdiag", "u1"), ("pad", "u1", (3,))])
assert ROW_DTYPE.itemsize == 48
#: compact record for meshes with at most 8 ring neighbours per vertex
ROW8_DTYPE = np.dtype([("g0", "<i4"), ("s0", "<u2"), ("nfc", "u1"), ("posdiag", "u1"),
                       ("nbr", "<u2", (8,)), ("pos", "u1", (8,))])
assert ROW8_DTYPE.itemsize == 32
MAXN = 12


class FanPlanError(Exception):
    """The mesh / pattern does not fit the vertex-fan kernel (caller falls back)."""


def _morton_vertices(p, bits=20):
    lo = p.min(axis=1, keepdims=True)
    ext = p.max(axis=1, keepdims=True) - lo
    ext[ext == 0] = 1.0
    q = np.minimum(((p - lo) / ext * (2 ** bits - 1)).astype(np.uint64),
                   np.uint64(2 ** bits - 1))
    key = np.zeros(p.shape[1], dtype=np.uint64)
    for b in range(bits):
        for d in range(p.shape[0]):
            key |= ((q[d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * p.shape[0] + d)
    return np.argsort(key, kind="stable")


def vertex_rings(mesh):
    """Counter-clockwise ring of neighbours around every vertex.

    Returns ``(nb, nfan, closed)``: ``nb[k, v]`` = k-th ring neighbour of vertex
    ``v`` (-1 beyond the ring), ``nfan[v]`` = number of triangles around ``v``,
    ``closed[v]`` = interior vertex (the ring closes).  Triangle ``k`` of the fan
    is ``(v, nb[k], nb[k+1])`` (indices modulo ``nfan`` for a closed ring)."""
    p, t = mesh.p, np.asarray(mesh.t)
    nt, nv = t.shape[1], p.shape[1]
    ar = np.arange(nt)
    x, y = p[0][t], p[1][t]
    det = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0])
    if np.any(det == 0):
        raise FanPlanError("degenerate triangle")
    ccw = det > 0
    li = np.arange(3)[:, None]
    nxt = np.where(ccw[None, :], (li + 1) % 3, (li + 2) % 3)     # CCW successor of local i
    prv = np.where(ccw[None, :], (li + 2) % 3, (li + 1) % 3)
    LF = np.array([[-1, 0, 2], [0, -1, 1], [2, 1, -1]])          # local facet of a vertex pair

    def across(i, j):
        fid = mesh.t2f[LF[i, j], ar]
        f0, f1 = mesh.f2t[0, fid], mesh.f2t[1, fid]
        return np.where(f0 == ar, f1, f0)

    next_corner = np.full(3 * nt, -1, dtype=np.int64)
    has_prev = np.zeros(3 * nt, dtype=bool)
    for i in range(3):
        ii = np.full(nt, i)
        r = t[i]
        tn = across(ii, prv[i])                  # next triangle CCW around r: across edge (r, b)
        ok = tn >= 0
        tnc = np.where(ok, tn, 0)
        iloc = np.argmax(t[:, tnc] == r[None, :], axis=0)
        next_corner[i * nt + ar] = np.where(ok, iloc * nt + tnc, -1)
        has_prev[i * nt + ar] = across(ii, nxt[i]) >= 0
    corner_vertex = t.reshape(-1)                # corner c = i*nt + T  ->  vertex
    valence = np.bincount(corner_vertex, minlength=nv)
    if valence.min() == 0:
        raise FanPlanError("vertex without elements")
    K = int(valence.max())
    if K > MAXN:
        raise FanPlanError("vertex valence %d > %d" % (K, MAXN))
    # start corner: the one without a predecessor (boundary), else the smallest
    cnxt = np.take(nxt.reshape(-1), np.arange(3 * nt))          # local index arrays per corner
    cprv = np.take(prv.reshape(-1), np.arange(3 * nt))
    corner_T = np.tile(ar, 3)
    a_of = t[cnxt, corner_T]                      # vertex following r (CCW) in the corner's triangle
    b_of = t[cprv, corner_T]
    # closed fans start at the neighbour with the smallest polar angle, so that
    # on a regular mesh the q-th neighbour of every vertex lies in the same
    # direction (neighbouring threads then read neighbouring coordinates)
    # (a monotone surrogate of atan2 without transcendentals, so that the native
    # builder -- host or device -- reproduces it bit for bit; ties: smallest corner)
    ang = angle_key(p[0][a_of] - p[0][corner_vertex], p[1][a_of] - p[1][corner_vertex])
    ordk = np.lexsort((ang, corner_vertex))
    start = ordk[np.searchsorted(corner_vertex[ordk], np.arange(nv))]
    nopred = np.nonzero(~has_prev)[0]
    start[corner_vertex[nopred]] = nopred
    nb = np.full((K + 1, nv), -1, dtype=np.int64)
    nfan = np.zeros(nv, dtype=np.int64)
    closed = np.zeros(nv, dtype=bool)
    cur = start.copy()
    alive = np.ones(nv, dtype=bool)
    last_b = np.full(nv, -1, dtype=np.int64)
    for k in range(K):
        idx = np.nonzero(alive)[0]
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
    if alive.any() or not np.array_equal(nfan, valence):
        raise FanPlanError("non-manifold vertex fan")
    opn = np.nonzero(~closed)[0]
    nb[nfan[opn], opn] = last_b[opn]              # open fans have one more neighbour
    return nb, nfan, closed


def angle_key(dx, dy):
    """Same ordering as ``np.arctan2(dy, dx)`` on (-pi, pi], values in (0, 4]."""
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(dy < 0,
                     np.where(dx < 0, (-dy) / (-dx - dy), 1.0 + dx / (dx - dy)),
                     np.where(dx > 0, 2.0 + dy / (dx + dy), 3.0 + (-dx) / (-dx + dy)))
    return np.where((dy == 0) & (dx < 0), 4.0, k)


def build_fan_plan(mesh, indptr, indices, rows_per_tile=256, tiny=True, row_order="id"):
    """Host arrays of the vertex-fan plan for the CSR pattern ``indptr/indices``
    (rows = columns = vertices).  Raises :class:`FanPlanError` if it does not
    apply."""
    p = mesh.p
    nv = p.shape[1]
    if p.shape[0] != 2 or len(indptr) != nv + 1:
        raise FanPlanError("not a scalar vertex-based pattern on a 2-D mesh")
    nb, nfan, closed = vertex_rings(mesh)
    nn = nfan + (~closed)
    if int(nn.max()) > MAXN:
        raise FanPlanError("more than %d ring neighbours" % MAXN)
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    rowlen = np.diff(indptr)
    if not np.array_equal(rowlen, nn + 1):
        raise FanPlanError("pattern rows are not vertex stars")
    if indptr[-1] > 2 ** 31 - 1:
        raise FanPlanError("nnz exceeds 32-bit positions")
    # CSR position (inside its row) of every ring neighbour and of the diagonal
    rowkey = np.repeat(np.arange(nv), rowlen) * nv + indices
    K1 = nb.shape[0]
    vv = np.broadcast_to(np.arange(nv), (K1, nv))
    valid = nb >= 0
    gpos = np.searchsorted(rowkey, vv[valid] * nv + nb[valid])
    if not np.array_equal(indices[np.minimum(gpos, len(indices) - 1)], nb[valid]):
        raise FanPlanError("ring neighbour missing from the pattern")
    pos = np.zeros((K1, nv), dtype=np.int64)
    pos[valid] = gpos - indptr[vv[valid]]
    dpos = np.searchsorted(rowkey, np.arange(nv) * nv + np.arange(nv)) - indptr[:-1]
    nb12 = np.full((MAXN, nv), -1, dtype=np.int64)
    pos12 = np.zeros((MAXN, nv), dtype=np.int64)
    kk = min(K1, MAXN)
    nb12[:kk], pos12[:kk] = nb[:kk], pos[:kk]
    # tiles of Morton-consecutive rows; inside a tile the rows follow mesh lines
    compact = int(nn.max()) <= 8
    rdt = ROW8_DTYPE if compact else ROW_DTYPE
    W = 8 if compact else MAXN
    R = int(rows_per_tile)
    ntiles = (nv + R - 1) // R
    vperm = _morton_vertices(p)
    tile_of_pos = np.arange(nv) // R
    if row_order == "id":
        order = np.lexsort((vperm, tile_of_pos))
    else:
        order = np.lexsort((p[0][vperm], p[1][vperm], tile_of_pos))
    rows = vperm[order]                                   # all rows, tile after tile
    tile_of_row = tile_of_pos                             # (positions are tile-major)
    nr = np.bincount(tile_of_row, minlength=ntiles)
    r0 = np.cumsum(nr) - nr
    row_local = np.arange(nv) - r0[tile_of_row]
    tile_of_vertex = np.empty(nv, dtype=np.int64)
    tile_of_vertex[rows] = tile_of_row
    local_of_vertex = np.empty(nv, dtype=np.int64)
    local_of_vertex[rows] = row_local
    # ring neighbours of every row, (W, nv) in row order
    nbr = nb12[:W, rows]
    ok = nbr >= 0
    nbr0 = np.where(ok, nbr, 0)
    tl = np.broadcast_to(tile_of_row, nbr.shape)
    inside = ok & (tile_of_vertex[nbr0] == tl)
    # halo vertices: unique (tile, vertex) pairs that are not rows of the tile
    hk = (tl[ok & ~inside] * nv + nbr[ok & ~inside])
    uh = np.unique(hk)
    h_tile = uh // nv
    nh = np.bincount(h_tile, minlength=ntiles)
    h0 = np.cumsum(nh) - nh
    nvt = nr + nh
    if int(nvt.max()) > 65535:
        raise FanPlanError("tile with more than 65535 vertices")
    loc = np.zeros(nbr.shape, dtype=np.int64)
    loc[inside] = local_of_vertex[nbr[inside]]
    out = ok & ~inside
    # halo vertices are numbered along mesh lines too (y, then x) so that the
    # q-th neighbours of consecutive rows are consecutive in shared memory
    hv = uh % nv
    hord = np.arange(len(uh))                     # halo vertices of a tile in id order
    hrank = np.empty(len(uh), dtype=np.int64)
    hrank[hord] = np.arange(len(uh))
    hidx = np.searchsorted(uh, tl[out] * nv + nbr[out])
    loc[out] = nr[tl[out]] + (hrank[hidx] - h0[tl[out]])
    # row records, as planes
    lens = rowlen[rows]
    ns = np.bincount(tile_of_row, weights=lens, minlength=ntiles).astype(np.int64)
    cl = np.cumsum(lens) - lens
    s0 = cl - cl[r0[tile_of_row]]
    head = np.zeros((nv, 2), dtype=np.int64)
    head[:, 0] = indptr[rows]                              # CSR position of the row
    head[:, 1] = (s0 | (nfan[rows] << 16) | (closed[rows].astype(np.int64) << 20)
                  | (dpos[rows] << 24))
    nb16 = loc.T.astype(np.uint16)                         # (nv, W) tile-local neighbour ids
    pos8 = np.where(ok, pos12[:W, rows], 0).T.astype(np.uint8)   # (nv, W) CSR positions
    # "tiny" records (20 bytes per row): 8-bit neighbour ids (tiles of <= 256
    # vertices) in plane 0, eight 4-bit CSR positions in one extra word
    tiny = bool(tiny) and compact and int(nvt.max()) <= 256
    plane0 = np.zeros((nv, 16), dtype=np.uint8)
    plane0[:, :8] = head.astype(np.int32).view(np.uint8).reshape(nv, 8)
    nib = None
    if tiny:
        plane0[:, 8:] = nb16[:, :8].astype(np.uint8)
        nib = np.zeros(nv, dtype=np.int64)
        for q in range(8):
            nib |= pos8[:, q].astype(np.int64) << (4 * q)
    else:
        plane0[:, 8:] = nb16[:, :4].copy().view(np.uint8).reshape(nv, 8)
    planes = [plane0]
    if tiny:
        pass
    elif compact:
        p1 = np.zeros((nv, 16), dtype=np.uint8)
        p1[:, :8] = nb16[:, 4:8].copy().view(np.uint8).reshape(nv, 8)
        p1[:, 8:] = pos8[:, :8]
        planes.append(p1)
    else:
        p1 = nb16[:, 4:12].copy().view(np.uint8).reshape(nv, 16)
        p2 = np.zeros((nv, 16), dtype=np.uint8)
        p2[:, :12] = pos8[:, :12]
        planes += [p1, p2]
    # blob layout
    bytes_a = (nvt * 16 + 15) // 16 * 16
    bytes_b = nr * 16 * len(planes)
    if tiny:
        bytes_b = bytes_b + (nr * 4 + 15) // 16 * 16
    tb = bytes_a + bytes_b
    off = np.cumsum(tb) - tb
    total = int(tb.sum())
    blob = np.zeros(total, dtype=np.uint8)
    for k, pl in enumerate(planes):
        dk = (off + bytes_a + k * nr * 16)[tile_of_row] + 16 * row_local
        blob[dk[:, None] + np.arange(16)[None, :]] = pl
    if tiny:
        dk = ((off + bytes_a + nr * 16)[tile_of_row] + 4 * row_local) // 4
        blob.view(np.uint32)[dk] = nib.astype(np.uint32)
    desc = np.zeros((ntiles, 8), dtype=np.int64)
    desc[:, 0], desc[:, 1], desc[:, 2] = off // 16, bytes_a, bytes_b
    desc[:, 3], desc[:, 4], desc[:, 5] = nr, ns, nvt
    if desc.max() > 2 ** 31 - 1:
        raise FanPlanError("fan plan offsets exceed 32 bits")
    # tile vertex lists: the rows first, then the halo vertices
    vid = np.concatenate((rows, hv))
    vtile = np.concatenate((tile_of_row, h_tile))
    vloc = np.concatenate((row_local, nr[h_tile] + hrank - h0[h_tile]))
    pc_index = off[vtile] // 8 + 2 * vloc
    return {
        "ntiles": ntiles, "rows_per_tile": R,
        "desc": desc.astype(np.int32), "blob": blob,
        "vid": vid, "pc_index": pc_index,
        "max_a": int(tb.max()), "max_b": 0, "max_ns": int(ns.max()),
        "halo_factor": float(nvt.sum()) / nv,
        "max_neighbours": int(nn.max()),
        "format": "tiny" if tiny else ("fan8" if compact else "wide"),
    }


def fill_coordinates_host(plan, p):
    """Host version of the coordinate refresh (tests)."""
    b64 = plan["blob"].view(np.float64)
    b64[plan["pc_index"]] = p[0][plan["vid"]]
    b64[plan["pc_index"] + 1] = p[1][plan["vid"]]
