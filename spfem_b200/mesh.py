# -*- coding: utf-8 -*-
"""Mesh containers: the *inputs* of the assembly hot path.

Drop-in for the classes of spfem/mesh.py that the hot path reads
(``MeshTri`` :773-1338, ``MeshTet`` :429-770, ``MeshQuad`` :222-426,
``MeshLine`` :154-219).  Attributes and numbering conventions are exactly the
reference's, because the global DOF numbering (``Dofnum``) and therefore the
CSR pattern are derived from them:

* ``p``  (dim x nv)  float64 vertices, ``t`` (nve x nt) int64 connectivity
  (triangles: every column sorted ascending, spfem/mesh.py:817-818),
* ``facets`` = unique sorted vertex tuples in lexicographic order,
  ``t2f`` / ``t2e`` element->facet / element->edge, ``edges`` (3-D only),
* ``f2t`` (2 x nf): first / last element seeing the facet in local-facet-major
  order, second row -1 on the boundary (spfem/mesh.py:837-851).

The topology is built with one shared routine (packed 64-bit keys + one sort)
instead of the reference's structured-view ``np.unique`` calls; the resulting
integer arrays are identical (the golden fixtures of the reference pin the DOF
numbering derived from them; ``tests/test_mesh_host.py`` checks the device
builder against this one).
Plotting / adaptive refinement / mesh IO are out of scope.
"""
import numpy as np


def _unique_columns(cols, nv):
    """Lexicographically ordered unique columns of the integer array ``cols``
    (k x n, entries in [0, nv)).  Returns ``(unique_cols, inverse)`` with the
    same ordering as ``np.unique`` on a structured view of the transposed
    array (spfem/mesh.py:466-470, 831-835)."""
    k = cols.shape[0]
    if k == 1 or float(nv) ** k < 2.0 ** 62:
        key = cols[0].astype(np.int64)
        for r in range(1, k):
            key = key * np.int64(nv) + cols[r]
        ukey, first, inverse = np.unique(key, return_index=True,
                                         return_inverse=True)
        return cols[:, first], inverse.reshape(-1)
    u, inverse = np.unique(cols.T, axis=0, return_inverse=True)
    return np.ascontiguousarray(u.T), inverse.reshape(-1)


def _facet_to_element(t2f, nf):
    """``f2t`` from ``t2f`` (rows = local facet index): first / last entry of
    each facet in the local-facet-major flattening (spfem/mesh.py:837-851)."""
    nloc, nt = t2f.shape
    flat = t2f.reshape(-1)
    order = np.argsort(flat, kind='stable')
    sorted_f = flat[order]
    start = np.searchsorted(sorted_f, np.arange(nf), side='left')
    stop = np.searchsorted(sorted_f, np.arange(nf), side='right') - 1
    f2t = np.zeros((2, nf), dtype=np.int64)
    f2t[0] = order[start] % nt
    f2t[1] = order[stop] % nt
    f2t[1, f2t[0] == f2t[1]] = -1
    return f2t


# ---------------------------------------------------------------------------
# device restatement of the two topology builders: same
# numbering as the NumPy versions above, bit for bit -- ``torch.unique`` on the
# packed vertex tuples is the lexicographic ``np.unique``, the stable argsort /
# searchsorted pair is the reference's first / last occurrence rule.  The
# tensors may live on the CPU (that is how tests/test_mesh_host.py checks them).
# ---------------------------------------------------------------------------
def _entity_table_torch(torch, t, local_tuples, nv):
    k = len(local_tuples[0])
    blocks = [torch.sort(t[list(loc), :], dim=0).values for loc in local_tuples]
    allc = torch.cat(blocks, 1)
    del blocks
    if float(nv) ** k >= 2.0 ** 62:
        # vertex triples of a large mesh (nv^3 > 2^62): two-level key -- the rank of
        # the leading pair among the distinct leading pairs is monotone in the
        # lexicographic order, so (rank, third vertex) sorts like the triple
        if k != 3 or float(nv) ** 2 >= 2.0 ** 62:
            raise ValueError("vertex tuples do not fit a 64-bit key")
        u01, r01 = torch.unique(allc[0] * nv + allc[1], return_inverse=True)
        if float(u01.numel()) * nv >= 2.0 ** 62:
            raise ValueError("vertex tuples do not fit a 64-bit key")
        key = r01 * nv + allc[2]
        del r01, allc
        ukey, inverse = torch.unique(key, return_inverse=True)
        del key
        pair = u01[ukey // nv]
        rows = [pair // nv, pair % nv, ukey % nv]
        return torch.stack(rows, 0), inverse.reshape(len(local_tuples), t.shape[1])
    key = allc[0].clone()
    for r in range(1, k):
        key = key * nv + allc[r]
    del allc
    ukey, inverse = torch.unique(key, return_inverse=True)
    del key
    rows = []
    for r in range(k):
        rows.append((ukey // (nv ** (k - 1 - r))) % nv)
    return torch.stack(rows, 0), inverse.reshape(len(local_tuples), t.shape[1])


def _facet_to_element_torch(torch, t2f, nf):
    nloc, nt = t2f.shape
    flat = t2f.reshape(-1)
    order = torch.argsort(flat, stable=True)
    sorted_f = flat[order]
    ar = torch.arange(nf, device=t2f.device)
    start = torch.searchsorted(sorted_f, ar, right=False)
    stop = torch.searchsorted(sorted_f, ar, right=True) - 1
    f0 = order[start] % nt
    f1 = order[stop] % nt
    f1 = torch.where(f0 == f1, torch.full_like(f1, -1), f1)
    return torch.stack((f0, f1), 0)


def _device_topology_wanted(nt):
    """Large meshes build their topology tables on the GPU when there is one
    (``SPFEM_DEVICE_TOPOLOGY=0`` keeps the NumPy path, ``=1`` forces the device
    path for any size)."""
    import os
    mode = os.environ.get("SPFEM_DEVICE_TOPOLOGY", "auto")
    if mode == "0" or (mode == "auto" and nt < (1 << 17)):
        return False
    try:
        import torch
    except ImportError:
        return False
    return torch.cuda.is_available()


class Mesh(object):
    """Finite element mesh (abstract base; spfem/mesh.py:41-152)."""

    refdom = "none"   #: reference domain of the elements
    brefdom = "none"  #: reference domain of the facets
    p = np.array([])
    t = np.array([])
    _local_facets = ()
    _local_edges = ()

    def __init__(self, p=None, t=None):
        pass

    def dim(self):
        """Spatial dimension -- a *float*, as in spfem/mesh.py:62-64."""
        return float(self.p.shape[0])

    def mapping(self):
        raise NotImplementedError("Default mapping not implemented!")

    def refine(self, N=1):
        """Perform one or more uniform refinements."""
        for _ in range(N):
            self._single_refine()

    def scale(self, scale):
        for itr in range(int(self.dim())):
            if isinstance(scale, tuple):
                self.p[itr, :] *= scale[itr]
            else:
                self.p[itr, :] *= scale

    def translate(self, vec):
        for itr in range(int(self.dim())):
            self.p[itr, :] += vec[itr]

    def remove_elements(self, element_indices):
        keep = np.setdiff1d(np.arange(self.t.shape[1]), element_indices)
        newt = self.t[:, keep]
        ptix = np.unique(newt)
        reverse = np.zeros(self.p.shape[1], dtype=np.int64)
        reverse[ptix] = np.arange(len(ptix))
        return self.p[:, ptix], reverse[newt].astype(np.intp)

    def _validate(self):
        """Mesh validity checks (spfem/mesh.py:120-152)."""
        if not np.issubdtype(np.asarray(self.t).dtype, np.integer):
            raise Exception("Mesh._validate(): Element connectivity "
                            "must consist of integers.")
        if self.p.shape[0] > 3:
            raise Exception("Mesh._validate(): We do not allow meshes "
                            "embedded into larger than 3-dimensional "
                            "Euclidean space! Please check that "
                            "the given vertex matrix is of size Ndim x Nvertices.")
        nvertices = {'line': 2, 'tri': 3, 'quad': 4, 'tet': 4}
        if self.t.shape[0] != nvertices[self.refdom]:
            raise Exception("Mesh._validate(): The given connectivity "
                            "matrix has wrong shape!")
        if self.p.shape[1] != np.unique(self.p.T, axis=0).shape[0]:
            raise Exception("Mesh._validate(): Mesh contains duplicate vertices.")
        if len(np.setdiff1d(np.arange(self.p.shape[1]), np.unique(self.t))):
            raise Exception("Mesh._validate(): Mesh contains a vertex "
                            "not belonging to any element.")

    # -- shared topology builder -------------------------------------------
    def _entity_table(self, local_tuples):
        """Unique sorted vertex tuples of the given local sub-entities and the
        element->entity table (rows follow ``local_tuples``)."""
        nt = self.t.shape[1]
        nv = self.p.shape[1]
        blocks = [np.sort(self.t[list(loc), :], axis=0) for loc in local_tuples]
        allc = np.hstack(blocks)
        uniq, inverse = _unique_columns(allc, max(nv, int(self.t.max()) + 1))
        return uniq, inverse.reshape((len(local_tuples), nt))

    def _build_mappings(self):
        if _device_topology_wanted(self.t.shape[1]):
            return self._build_mappings_device()
        if self._local_edges:
            self.edges, self.t2e = self._entity_table(self._local_edges)
        self.facets, self.t2f = self._entity_table(self._local_facets)
        self.f2t = _facet_to_element(self.t2f, self.facets.shape[1])

    def _build_mappings_device(self, device=None):
        """The same tables built with sort / unique on the GPU and copied back
        (the mesh classes keep NumPy arrays, like the reference)."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        t = torch.from_numpy(np.ascontiguousarray(self.t, dtype=np.int64)).to(dev)
        nv = max(self.p.shape[1], int(t.max().item()) + 1)
        if self._local_edges:
            e, t2e = _entity_table_torch(torch, t, self._local_edges, nv)
            self.edges, self.t2e = e.cpu().numpy(), t2e.cpu().numpy()
        f, t2f = _entity_table_torch(torch, t, self._local_facets, nv)
        f2t = _facet_to_element_torch(torch, t2f, f.shape[1])
        self.facets, self.t2f, self.f2t = f.cpu().numpy(), t2f.cpu().numpy(), f2t.cpu().numpy()

    # -- queries shared by the 2-D / 3-D meshes -----------------------------
    def boundary_facets(self):
        return np.nonzero(self.f2t[1, :] == -1)[0]

    def interior_facets(self):
        return np.nonzero(self.f2t[1, :] >= 0)[0]

    def boundary_nodes(self):
        return np.unique(self.facets[:, self.boundary_facets()])

    def interior_nodes(self):
        return np.setdiff1d(np.arange(0, self.p.shape[1]),
                            self.boundary_nodes())

    def nodes_satisfying(self, test):
        return np.nonzero(test(*[self.p[k, :] for k in range(self.p.shape[0])]))[0]

    def facets_satisfying(self, test):
        """Facets whose midpoints satisfy ``test``."""
        nfv = self.facets.shape[0]
        w = 0.5 if nfv == 2 else 0.3333333  # sic: spfem/mesh.py:520-528
        mids = [w * sum(self.p[k, self.facets[j, :]] for j in range(nfv))
                for k in range(self.p.shape[0])]
        return np.nonzero(test(*mids))[0]

    def param(self):
        """Mesh parameter: the longest edge."""
        ed = self.edges if self._local_edges else self.facets
        return np.max(np.sqrt(np.sum((self.p[:, ed[0, :]] -
                                      self.p[:, ed[1, :]]) ** 2, axis=0)))


class MeshLine(Mesh):
    """One-dimensional mesh (spfem/mesh.py:154-219). Not on the hot path."""

    refdom = "line"
    brefdom = "point"

    def __init__(self, p=None, t=None, validate=True):
        if p is None and t is None:
            p = np.array([[0, 1]], dtype=np.float64)
            t = np.array([[0], [1]], dtype=np.int64)
        elif p is None or t is None:
            raise Exception("Must provide p AND t or neither")
        self.p = p
        self.t = t
        if validate:
            self._validate()

    def _single_refine(self):
        t, p = self.t, self.p
        mid = np.arange(t.shape[1]) + np.max(t) + 1
        self.p = np.hstack((p, 0.5 * (p[:, t[0, :]] + p[:, t[1, :]])))
        self.t = np.hstack((np.vstack((t[0, :], mid)),
                            np.vstack((mid, t[1, :]))))

    def boundary_nodes(self):
        _, counts = np.unique(self.t.flatten(), return_counts=True)
        return np.nonzero(counts == 1)[0]

    def interior_nodes(self):
        _, counts = np.unique(self.t.flatten(), return_counts=True)
        return np.nonzero(counts == 2)[0]

    def mapping(self):
        from . import mapping
        return mapping.MappingAffine(self)


class MeshQuad(Mesh):
    """Quadrilateral mesh (spfem/mesh.py:222-426); vertices of every element
    are kept in the given counter-clockwise order (no sorting, :241-243)."""

    refdom = "quad"
    brefdom = "line"
    _local_facets = ((0, 1), (1, 2), (2, 3), (0, 3))

    def __init__(self, p=None, t=None, validate=True):
        if p is None and t is None:
            p = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float64).T
            t = np.array([[0, 1, 2, 3]], dtype=np.int64).T
        elif p is None or t is None:
            raise Exception("Must provide p AND t or neither")
        self.p = p
        self.t = t
        if validate:
            self._validate()
        self._build_mappings()

    def _single_refine(self):
        """Split every quad into four (spfem/mesh.py:312-357)."""
        t, p, e = self.t, self.p, self.facets
        sz = p.shape[1]
        t2f = self.t2f + sz
        mid = np.arange(t.shape[1]) + np.max(t2f) + 1
        newp1 = 0.5 * np.vstack((p[0, e[0, :]] + p[0, e[1, :]],
                                 p[1, e[0, :]] + p[1, e[1, :]]))
        newp2 = 0.25 * np.vstack((p[0, t[0, :]] + p[0, t[1, :]] +
                                  p[0, t[2, :]] + p[0, t[3, :]],
                                  p[1, t[0, :]] + p[1, t[1, :]] +
                                  p[1, t[2, :]] + p[1, t[3, :]]))
        self.p = np.hstack((p, newp1, newp2))
        self.t = np.hstack((np.vstack((t[0, :], t2f[0, :], mid, t2f[3, :])),
                            np.vstack((t2f[0, :], t[1, :], t2f[1, :], mid)),
                            np.vstack((mid, t2f[1, :], t[2, :], t2f[2, :])),
                            np.vstack((t2f[3, :], mid, t2f[2, :], t[3, :]))))
        self._build_mappings()

    def _splitquads(self, x):
        if len(x) == self.t.shape[1]:
            X = np.concatenate((x, x))
        else:
            X = x
        t = np.hstack((self.t[[0, 1, 3], :], self.t[[1, 2, 3]]))
        return MeshTri(self.p, t), X

    def jiggle(self, z=0.2):
        y = z * self.param()
        I = self.interior_nodes()
        self.p[0, I] = self.p[0, I] + y * np.random.rand(len(I))
        self.p[1, I] = self.p[1, I] + y * np.random.rand(len(I))

    def mapping(self):
        from . import mapping
        return mapping.MappingQ1(self)


class MeshTet(Mesh):
    """Tetrahedral mesh (spfem/mesh.py:429-770).  Local edge order
    (0,1)(1,2)(0,2)(0,3)(1,3)(2,3), local facets (0,1,2)(0,1,3)(0,2,3)(1,2,3)
    (:451-483)."""

    refdom = "tet"
    brefdom = "tri"
    _local_edges = ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))
    _local_facets = ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))

    def __init__(self, p=None, t=None, validate=True):
        if p is None and t is None:
            p = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0],
                          [0, 1, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1]],
                         dtype=np.float64).T
            t = np.array([[0, 1, 2, 3], [3, 5, 1, 7], [2, 3, 6, 7],
                          [2, 3, 1, 7], [1, 2, 4, 7]], dtype=np.int64).T
        elif p is None or t is None:
            raise Exception("Must provide p AND t or neither")
        self.p = p
        self.t = t
        if validate:
            self._validate()
        self._build_mappings()

    def edges_satisfying(self, test):
        mids = [0.5 * (self.p[k, self.edges[0, :]] + self.p[k, self.edges[1, :]])
                for k in range(3)]
        return np.nonzero(test(*mids))[0]

    def boundary_edges(self):
        onb = np.zeros(self.p.shape[1], dtype=bool)
        onb[self.boundary_nodes()] = True
        return np.nonzero(onb[self.edges[0, :]] & onb[self.edges[1, :]])[0]

    def shapereg(self):
        def edgelen(n):
            return np.sqrt(np.sum((self.p[:, self.edges[0, self.t2e[n, :]]] -
                                   self.p[:, self.edges[1, self.t2e[n, :]]]) ** 2,
                                  axis=0))
        edgelenmat = np.vstack(tuple(edgelen(i) for i in range(6)))
        return np.max(np.max(edgelenmat, axis=0) / np.min(edgelenmat, axis=0))

    def _single_refine(self):
        """Red refinement into 8 sub-tetrahedra (spfem/mesh.py:539-631): four
        corner tets, then the inner octahedron is split along its shortest
        diagonal -- as in the reference, 'shortest' is measured with the x and
        y coordinates only (:584-589)."""
        t, p, e = self.t, self.p, self.edges
        sz = p.shape[1]
        m = self.t2e + sz        # midpoint vertex of local edge k = m[k]
        newp = np.hstack((p, 0.5 * (p[:, e[0, :]] + p[:, e[1, :]])))

        def sqdist_xy(a, b):
            return ((newp[0, m[a]] - newp[0, m[b]]) ** 2 +
                    (newp[1, m[a]] - newp[1, m[b]]) ** 2)
        d1, d2, d3 = sqdist_xy(2, 4), sqdist_xy(1, 3), sqdist_xy(0, 5)
        c1 = (d1 < d2) & (d1 < d3)
        c2 = (~(d1 < d2)) & (d2 < d3)
        c3 = (~(d1 < d3)) & (~(d2 < d3))
        blocks = [np.vstack((t[0], m[0], m[2], m[3])),
                  np.vstack((t[1], m[0], m[1], m[4])),
                  np.vstack((t[2], m[1], m[2], m[5])),
                  np.vstack((t[3], m[3], m[4], m[5]))]
        inner = ((c1, (2, 4), ((0, 1), (0, 3), (1, 5), (3, 5))),
                 (c2, (1, 3), ((0, 4), (4, 5), (5, 2), (2, 0))),
                 (c3, (0, 5), ((1, 4), (4, 3), (3, 2), (2, 1))))
        for mask, (da, db), others in inner:
            for (oa, ob) in others:
                blocks.append(np.vstack((m[da, mask], m[db, mask],
                                         m[oa, mask], m[ob, mask])))
        self.p = newp
        self.t = np.hstack(blocks)
        self._build_mappings()

    def mapping(self):
        from . import mapping
        return mapping.MappingAffine(self)


class MeshTri(Mesh):
    """Triangular mesh (spfem/mesh.py:773-1338).  Local facets
    (0,1)(1,2)(0,2) (:820-827)."""

    refdom = "tri"
    brefdom = "line"
    _local_facets = ((0, 1), (1, 2), (0, 2))

    def __init__(self, p=None, t=None, validate=True, initmesh=None,
                 sort_t=True):
        if p is None and t is None:
            if initmesh == 'symmetric':
                p = np.array([[0, 1, 1, 0, 0.5],
                              [0, 0, 1, 1, 0.5]], dtype=np.float64)
                t = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [0, 3, 4]],
                             dtype=np.intp).T
            elif initmesh == 'sqsymmetric':
                p = np.array([[0, 0.5, 1, 0, 0.5, 1, 0, 0.5, 1],
                              [0, 0, 0, 0.5, 0.5, 0.5, 1, 1, 1]],
                             dtype=np.float64)
                t = np.array([[0, 1, 4], [1, 2, 4], [2, 4, 5], [0, 3, 4],
                              [3, 4, 6], [4, 6, 7], [4, 7, 8], [4, 5, 8]],
                             dtype=np.intp).T
            elif initmesh == 'reftri':
                p = np.array([[0, 1, 0], [0, 0, 1]], dtype=np.float64)
                t = np.array([[0, 1, 2]], dtype=np.intp).T
            else:
                p = np.array([[0, 1, 0, 1], [0, 0, 1, 1]], dtype=np.float64)
                t = np.array([[0, 1, 2], [1, 3, 2]], dtype=np.intp).T
        elif p is None or t is None:
            raise Exception("Must provide p AND t or neither")
        self.p = p
        self.t = t
        if validate:
            self._validate()
        self._build_mappings(sort_t=sort_t)

    def _build_mappings(self, sort_t=True):
        if sort_t:
            # ascending vertex ids per triangle: orientation (sign of detA)
            # becomes mesh dependent -- |detDF| in the assembly is essential
            self.t = np.sort(self.t, axis=0)
        Mesh._build_mappings(self)

    def elements_satisfying(self, test):
        mx = .33333 * np.sum(self.p[0, self.t], axis=0)
        my = .33333 * np.sum(self.p[1, self.t], axis=0)
        return np.nonzero(test(mx, my))[0]

    def _single_refine(self):
        """Red refinement into 4 triangles (spfem/mesh.py:1045-1066)."""
        t, p, e = self.t, self.p, self.facets
        sz = p.shape[1]
        m = self.t2f + sz
        self.p = np.hstack((p, 0.5 * np.vstack((p[0, e[0, :]] + p[0, e[1, :]],
                                                p[1, e[0, :]] + p[1, e[1, :]]))))
        self.t = np.hstack((np.vstack((t[0], m[0], m[2])),
                            np.vstack((t[1], m[0], m[1])),
                            np.vstack((t[2], m[2], m[1])),
                            np.vstack((m[0], m[1], m[2]))))
        self._build_mappings()

    def _neighbors(self, tris=None):
        if tris is None:
            tris = np.arange(self.t.shape[1])
        rows = []
        for k in range(3):
            ft = self.f2t[:, self.t2f[k, tris]]
            rows.append(ft[ft != tris].flatten('F'))
        neighbors = np.vstack(rows)
        ix = (neighbors == -1)
        neighbors[ix] = np.vstack((tris, tris, tris))[ix]
        return neighbors

    def smooth(self, c=1.0):
        """Laplacian smoothing of the interior nodes (spfem/mesh.py:908-925;
        the reference body uses an undefined name ``k``; ``c`` is meant)."""
        from .assembly import AssemblerElement
        from .element import ElementTriP1
        from .utils import direct
        a = AssemblerElement(self, ElementTriP1())
        K = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
        M = a.iasm(lambda u, v: u * v)
        I = self.interior_nodes()
        dx = -c * direct(M, K.dot(self.p[0, :]))
        dy = -c * direct(M, K.dot(self.p[1, :]))
        self.p[0, I] += dx[I]
        self.p[1, I] += dy[I]

    def mapping(self):
        from . import mapping
        return mapping.MappingAffine(self)
