# -*- coding: utf-8 -*-
"""Structured synthetic meshes for the benchmark configurations of SURVEY.md
section 8(d): an ``n x n`` grid of the unit square cut into ``2 n^2`` triangles
(config 3 / 5: n = 3162 -> 19 996 488 triangles) and the Kuhn (Freudenthal)
triangulation of an ``m^3`` grid of the unit cube into ``6 m^3`` tetrahedra
(config 4: m = 203 -> 50 193 162 tetrahedra).  The reference has no structured
generator (its meshes come from ``refine`` or meshpy); these only produce the
``(p, t)`` arrays that ``MeshTri(p, t)`` / ``MeshTet(p, t)`` take, optionally for
a sub-box of cells so that a rank of the multi-GPU run can build its own piece
(plus a halo layer of cells) without ever materialising the global mesh."""
import itertools

import numpy as np


def grid_tri(n, m=None, i0=0, i1=None, j0=0, j1=None, nx=None, ny=None):
    """``(p, t)`` of the cells ``[i0, i1) x [j0, j1)`` of an ``nx x ny`` grid of the
    rectangle ``[0, nx/n] x [0, ny/n]`` (default: the unit square, all cells).
    Vertices are numbered row by row inside the sub-box; every cell gives the two
    triangles ``(a, b, c), (b, d, c)`` of the reference's default mesh
    (spfem/mesh.py:806-811)."""
    nx = n if nx is None else nx
    ny = (n if m is None else m) if ny is None else ny
    i1 = nx if i1 is None else i1
    j1 = ny if j1 is None else j1
    ci, cj = i1 - i0, j1 - j0
    xs = np.arange(i0, i1 + 1, dtype=np.float64) / n
    ys = np.arange(j0, j1 + 1, dtype=np.float64) / n
    X, Y = np.meshgrid(xs, ys)                       # (cj+1, ci+1), row j
    p = np.vstack((X.reshape(-1), Y.reshape(-1)))
    I, J = np.meshgrid(np.arange(ci, dtype=np.int64), np.arange(cj, dtype=np.int64))
    a = (J * (ci + 1) + I).reshape(-1)
    b, c, d = a + 1, a + (ci + 1), a + (ci + 2)
    t = np.empty((3, 2 * ci * cj), dtype=np.int64)
    t[:, 0::2] = np.vstack((a, b, c))
    t[:, 1::2] = np.vstack((b, c, d))                # (b, d, c) with ascending ids
    return p, t


def grid_tet(m, k0=0, k1=None, mx=None, my=None, mz=None):
    """``(p, t)`` of the Kuhn triangulation of the cell layers ``[k0, k1)`` (along
    z) of an ``mx x my x mz`` grid with cell size ``1/m``: six tetrahedra per cell
    around the main diagonal, conforming across cells."""
    mx = m if mx is None else mx
    my = m if my is None else my
    mz = m if mz is None else mz
    k1 = mz if k1 is None else k1
    ck = k1 - k0
    xs = np.arange(mx + 1, dtype=np.float64) / m
    ys = np.arange(my + 1, dtype=np.float64) / m
    zs = np.arange(k0, k1 + 1, dtype=np.float64) / m
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    p = np.vstack((X.reshape(-1), Y.reshape(-1), Z.reshape(-1)))
    sx, sy, sz = 1, mx + 1, (mx + 1) * (my + 1)
    K, J, I = np.meshgrid(np.arange(ck, dtype=np.int64), np.arange(my, dtype=np.int64),
                          np.arange(mx, dtype=np.int64), indexing="ij")
    v0 = (K * sz + J * sy + I * sx).reshape(-1)
    step = (sx, sy, sz)
    nc = v0.shape[0]
    t = np.empty((4, 6 * nc), dtype=np.int64)
    for q, perm in enumerate(itertools.permutations(range(3))):
        v1 = v0 + step[perm[0]]
        v2 = v1 + step[perm[1]]
        v3 = v2 + step[perm[2]]
        t[:, q::6] = np.vstack((v0, v1, v2, v3))
    return p, t
