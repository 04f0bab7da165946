# -*- coding: utf-8 -*-
"""Multi-GPU assembly: one process per GPU, element-partitioned mesh.

The reference has no parallelism at all; this is the
B200 design of BASELINE.json's north star.  Elements are put in Morton order
and cut into ``world`` contiguous blocks.  Every global DOF row is *owned* by
the lowest rank whose block contains an element touching it.  A rank assembles
its block **plus the halo elements that touch rows it owns** (an O(surface)
layer that the neighbouring rank computes as well), so every owned row is
complete after the rank's own fused kernel:

* no data-path collective, no partial sums crossing NVLink (``mode="halo"``,
  the default; ``mode="exchange"`` is the north star's alternative: own block
  only, ghost-row partial sums sent to the owners in one ``all_to_all`` and
  added in fixed rank order, see :class:`InterfaceExchange`),
* the summation order inside a CSR slot is fixed by the local contribution
  lists => bitwise reproducible for a given partition,
* the global ``indptr/indices`` are exactly the single-GPU pattern: ownership
  only partitions rows, DOF numbering is never changed.

The result stays row-distributed on the devices (:class:`DistResult`); it is
gathered into one host SciPy object only on request (``gather``), through
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

``partition_global`` splits a replicated host mesh; ``strip_part`` builds the
rank's piece of the weak-scaling benchmark mesh (a strip of ``world`` unit
squares, each ``MeshTri().refine(R)``) without ever materialising the global
mesh.
"""
import numpy as np
from scipy.sparse import coo_matrix

from . import mesh as _mesh
from .assembly import AssemblerElement, Dofnum


def morton_order_host(mesh, bits=20):
    """Permutation of the elements along a Morton curve through the centroids
    (host restatement of ``spfem_morton_order``; partitioning is set-up work)."""
    dim = mesh.p.shape[0]
    c = mesh.p[:, mesh.t].mean(axis=1)
    lo = mesh.p.min(axis=1, keepdims=True)
    ext = mesh.p.max(axis=1, keepdims=True) - lo
    ext[ext == 0] = 1.0
    q = np.minimum(((c - lo) / ext * (2 ** bits - 1)).astype(np.uint64),
                   np.uint64(2 ** bits - 1))
    key = np.zeros(c.shape[1], dtype=np.uint64)
    for b in range(bits):
        for d in range(dim):
            key |= ((q[d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * dim + d)
    return np.argsort(key, kind="stable")


class LocalPart(object):
    """A rank's share of the mesh: the local mesh (own block + halo), the maps
    from local to global DOFs and the mask of owned test DOFs (rows)."""

    def __init__(self, mesh, l2g_u, l2g_v, owned_v, N_u, N_v, n_own_elements,
                 owner_v=None):
        self.mesh, self.l2g_u, self.l2g_v = mesh, l2g_u, l2g_v
        self.owned_v, self.N_u, self.N_v = owned_v, N_u, N_v
        self.n_own_elements = n_own_elements
        # owning rank of every local test DOF (needed by the interface exchange)
        self.owner_v = owner_v
        # local facet ids -> bool "is a boundary facet of the GLOBAL mesh" (the
        # local mesh also has artificial boundary facets along the cuts)
        self.on_global_boundary = None

    def global_boundary_facets(self):
        """Local ids of the facets that lie on the boundary of the global mesh."""
        lb = self.mesh.boundary_facets()
        if self.on_global_boundary is None:
            raise NotImplementedError("this partition does not know the global boundary")
        return lb[self.on_global_boundary(self.mesh, lb)]


class Partitioner(object):
    """Element partition of a replicated host mesh into ``world`` Morton blocks.
    The rank-independent work (Morton order, global DOF numbering, row owners)
    is done once; :meth:`part` cuts out one rank's :class:`LocalPart` (the
    chunked single-GPU assembly asks for every part in turn)."""

    def __init__(self, mesh, elem_u, elem_v, world):
        self.mesh, self.elem_u, self.elem_v, self.world = mesh, elem_u, elem_v, world
        nt = mesh.t.shape[1]
        perm = morton_order_host(mesh)
        self.rank_of_elem = np.empty(nt, dtype=np.int64)
        bounds = [(nt * r) // world for r in range(world + 1)]
        for r in range(world):
            self.rank_of_elem[perm[bounds[r]:bounds[r + 1]]] = r
        self.dn_u = Dofnum(mesh, elem_u)
        self.dn_v = (self.dn_u if elem_v is None or elem_v is elem_u
                     else Dofnum(mesh, elem_v))
        # a row belongs to the lowest rank touching it: write the ranks in
        # descending order, the lowest one last
        self.owner_v = np.full(self.dn_v.N, world, dtype=np.int64)
        for r in range(world - 1, -1, -1):
            self.owner_v[self.dn_v.t_dof[:, self.rank_of_elem == r]] = r
        self._gb = None

    def _global_boundary(self):
        if self._gb is None:
            m = self.mesh
            self._gb = np.sort(m.facets[:, m.boundary_facets()], axis=0)
        return self._gb

    def part(self, rank, halo=True, facet_halo=False):
        """``halo=True``: own block + the elements touching owned rows (owned
        rows are complete locally).  ``halo=False``: the own block only; rows
        that the rank touches but does not own carry partial sums that
        :class:`InterfaceExchange` sends to their owners.  ``facet_halo=True``
        (with ``halo``) adds the facet neighbours of those elements: an interior
        facet contributes to the rows of the elements on *both* its sides, so
        interior-facet forms need the element across every facet of an element
        that touches an owned row."""
        mesh, dn_u, dn_v, owner_v = self.mesh, self.dn_u, self.dn_v, self.owner_v
        if halo:
            mine = (owner_v[dn_v.t_dof] == rank).any(axis=0)
            if facet_halo:
                nb = mesh.f2t[:, mesh.t2f[:, mine].reshape(-1)].reshape(-1)
                mine = mine.copy()
                mine[nb[nb >= 0]] = True
        else:
            if facet_halo:
                raise NotImplementedError("facet_halo needs halo=True")
            mine = self.rank_of_elem == rank
        elems = np.nonzero(mine)[0]
        verts = np.unique(mesh.t[:, elems])
        t_loc = np.searchsorted(verts, mesh.t[:, elems])
        local = type(mesh)(np.ascontiguousarray(mesh.p[:, verts]), t_loc, validate=False)
        ldn_u = Dofnum(local, self.elem_u)
        ldn_v = ldn_u if dn_v is dn_u else Dofnum(local, self.elem_v)
        l2g_u = np.empty(ldn_u.N, dtype=np.int64)
        l2g_u[ldn_u.t_dof] = dn_u.t_dof[:, elems]
        l2g_v = l2g_u
        if ldn_v is not ldn_u:
            l2g_v = np.empty(ldn_v.N, dtype=np.int64)
            l2g_v[ldn_v.t_dof] = dn_v.t_dof[:, elems]
        owned = owner_v[l2g_v] == rank
        part = LocalPart(local, l2g_u, l2g_v, owned, dn_u.N, dn_v.N,
                         int(np.sum(self.rank_of_elem[elems] == rank)),
                         owner_v=owner_v[l2g_v])

        def on_global_boundary(lmesh, lf):
            """Match the (sorted, global) vertex tuples of local facets ``lf``
            against the global boundary facets."""
            gb = self._global_boundary()
            mine = np.sort(verts[lmesh.facets[:, lf]], axis=0)
            both = np.hstack((gb, mine)).T
            _, inv = np.unique(both, axis=0, return_inverse=True)
            inv = inv.reshape(-1)
            hit = np.zeros(inv.max() + 1, dtype=bool)
            hit[inv[:gb.shape[1]]] = True
            return hit[inv[gb.shape[1]:]]

        part.on_global_boundary = on_global_boundary
        return part


def partition_global(mesh, elem_u, elem_v, world, rank, halo=True, facet_halo=False):
    """Split the replicated global ``mesh`` for ``rank`` of ``world``
    (see :class:`Partitioner`)."""
    return Partitioner(mesh, elem_u, elem_v, world).part(rank, halo=halo,
                                                         facet_halo=facet_halo)


def strip_part(rank, world, refine, unit=None, halo=True):
    """Rank ``rank``'s part of the weak-scaling mesh: ``world`` copies of the
    unit-square mesh ``MeshTri().refine(refine)`` glued side by side along x
    (copy r = unit mesh translated by (r, 0); the vertices on x = r+1 are
    shared by copies r and r+1).  ElementTriP1 DOFs (vertices).

    Global vertex numbering: copy r contributes its vertices except (r > 0)
    those on its left edge, in unit-mesh order; vertex v of copy r gets id
    ``r*(nv - nb) + c(v) + (nb if r > 0 ... )`` -- see ``gid`` below.  Ownership
    follows the general rule (lowest rank touching the row): rank r owns every
    vertex of its copy except its left edge.  The halo is the layer of copy
    r+1's elements touching x = r+1; ``halo=False`` leaves it out (the part for
    ``DistributedAssembler(mode="exchange")``: the left-edge rows are ghost
    rows whose partial sums go to rank r-1)."""
    if unit is None:
        unit = _mesh.MeshTri()
        unit.refine(refine)
    p, t = unit.p, unit.t
    nv = p.shape[1]
    left = np.nonzero(p[0] == 0.0)[0]
    right = np.nonzero(p[0] == 1.0)[0]
    # match left-edge and right-edge vertices by their y coordinate
    lo = left[np.argsort(p[1, left], kind="stable")]
    ro = right[np.argsort(p[1, right], kind="stable")]
    assert np.array_equal(p[1, lo], p[1, ro])
    nb = len(left)
    is_left = np.zeros(nv, dtype=bool)
    is_left[left] = True
    # compact index of the non-left vertices of one copy
    cidx = np.cumsum(~is_left) - 1
    partner = np.full(nv, -1, dtype=np.int64)      # left vertex -> right vertex
    partner[lo] = ro

    def gid(copy, v):
        """Global id of unit vertex ``v`` (array) in copy ``copy``."""
        v = np.asarray(v)
        out = np.empty(v.shape, dtype=np.int64)
        if copy == 0:
            return v.astype(np.int64)
        base = nv + (copy - 1) * (nv - nb)
        lf = is_left[v]
        out[~lf] = base + cidx[v[~lf]]
        out[lf] = gid(copy - 1, partner[v[lf]])
        return out

    N = nv + (world - 1) * (nv - nb)
    # own elements + halo (copy rank+1's elements touching its left edge)
    halo = (np.nonzero(is_left[t].any(axis=0))[0] if rank < world - 1 and halo
            else np.zeros(0, np.int64))
    hv = np.unique(t[:, halo])
    hv_new = hv[~is_left[hv]]                     # halo vertices not on the shared edge
    # local numbering: unit vertices 0..nv-1, then the new halo vertices
    hmap = np.full(nv, -1, dtype=np.int64)
    hmap[hv_new] = nv + np.arange(len(hv_new))
    hmap[lo] = ro                                  # shared edge: my right-edge vertices
    p_loc = np.hstack((p + np.array([[float(rank)], [0.0]]),
                       p[:, hv_new] + np.array([[float(rank + 1)], [0.0]])))
    t_loc = np.hstack((t, hmap[t[:, halo]]))
    local = _mesh.MeshTri(np.ascontiguousarray(p_loc), np.ascontiguousarray(t_loc),
                          validate=False)
    l2g = np.concatenate((gid(rank, np.arange(nv)), gid(rank + 1, hv_new)
                          if len(hv_new) else np.zeros(0, np.int64)))
    owned = np.zeros(p_loc.shape[1], dtype=bool)
    owned[:nv] = True
    owner = np.full(p_loc.shape[1], rank + 1, dtype=np.int64)   # new halo vertices
    owner[:nv] = rank
    if rank > 0:
        owned[left] = False
        owner[left] = rank - 1
    part = LocalPart(local, l2g, l2g, owned, N, N, t.shape[1], owner_v=owner)

    def on_global_boundary(lmesh, lf):
        """A local boundary facet is on the strip's boundary iff its midpoint is
        on y = 0, y = 1, x = 0 or x = world (the rest are cuts)."""
        mid = lmesh.p[:, lmesh.facets[:, lf]].mean(axis=1)
        return ((mid[1] == 0.0) | (mid[1] == 1.0) | (mid[0] == 0.0)
                | (mid[0] == float(world)))

    part.on_global_boundary = on_global_boundary
    return part


class SlabPart(LocalPart):
    """Rank ``rank``'s piece of a structured grid cut into ``world`` slabs along
    the last axis (:func:`slab_part`).  The maps to global DOF numbers are built
    on first use: analytically for vertex DOFs and for the edges of the 2-D grid,
    otherwise by matching vertex tuples against the global mesh (small meshes:
    tests and the benchmark's parity check)."""

    def __init__(self, mesh, elem_u, elem_v, owned_v, owner_v, n_own_elements, info):
        self.mesh, self.owned_v, self.owner_v = mesh, owned_v, owner_v
        self.n_own_elements = n_own_elements
        self.elem_u, self.elem_v, self.info = elem_u, elem_v, info
        self.on_global_boundary = None
        self._maps = None

    def _global_mesh(self):
        from . import meshgen
        i = self.info
        if i["n_total"] > 4000000:
            raise NotImplementedError("global DOF ids of this element on a slab "
                                      "partition need the global mesh (small meshes only)")
        if i["kind"] == "tri":
            p, t = meshgen.grid_tri(i["size"], ny=i["size"] * i["world"])
            return _mesh.MeshTri(p, t, validate=False)
        p, t = meshgen.grid_tet(i["size"], mz=i["size"] * i["world"])
        return _mesh.MeshTet(p, t, validate=False)

    @staticmethod
    def _tri_facet_index(v0, v1, nx, ny):
        """Index of the edge (v0 < v1) of the nx x ny triangle grid in the
        lexicographic list of its sorted vertex pairs (= ``mesh.facets``)."""
        s = nx + 1
        I, J = v0 % s, v0 // s
        full_rows = J * (3 * nx + 1)
        prefix = np.where(J < ny, np.where(I == 0, 0, 3 * I - 1), I)
        d = v1 - v0
        pos = np.where(d == 1, 0, np.where(d == nx, (I < nx).astype(np.int64),
                                           (I < nx).astype(np.int64) + (I >= 1)))
        return full_rows + prefix + pos

    def _dof_maps(self):
        if self._maps is not None:
            return self._maps
        i = self.info
        lm = self.mesh
        voff = i["vertex_offset"]
        maps = []
        gm = None
        for elem in (self.elem_u, self.elem_v):
            if elem is None or (maps and elem is self.elem_u):
                maps.append(maps[0])
                continue
            ldn = Dofnum(lm, elem)
            l2g = np.empty(ldn.N, dtype=np.int64)
            nvg = i["nv_global"]
            offset = elem.n_dofs * nvg
            if elem.n_dofs:
                k = np.arange(elem.n_dofs)[:, None]
                l2g[ldn.n_dof] = (np.arange(lm.p.shape[1])[None, :] + voff) * elem.n_dofs + k
            N = offset
            analytic = i["kind"] == "tri"
            if not analytic and (elem.e_dofs or elem.f_dofs or elem.i_dofs):
                gm = gm or self._global_mesh()
            if hasattr(lm, "edges"):
                ne_g = gm.edges.shape[1] if gm is not None else 0
                if elem.e_dofs:
                    gi = _match_tuples(lm.edges + voff, gm.edges, nvg)
                    l2g[ldn.e_dof] = offset + gi[None, :] * elem.e_dofs + np.arange(elem.e_dofs)[:, None]
                offset += elem.e_dofs * ne_g
            if analytic:
                nx, ny = i["size"], i["size"] * i["world"]
                nf_g = nx * (ny + 1) + (nx + 1) * ny + nx * ny
                gi = (self._tri_facet_index(lm.facets[0] + voff, lm.facets[1] + voff, nx, ny)
                      if elem.f_dofs else None)
            else:
                nf_g = gm.facets.shape[1] if gm is not None else 0
                gi = _match_tuples(lm.facets + voff, gm.facets, nvg) if elem.f_dofs else None
            if elem.f_dofs:
                l2g[ldn.f_dof] = offset + gi[None, :] * elem.f_dofs + np.arange(elem.f_dofs)[:, None]
            offset += elem.f_dofs * nf_g
            if elem.i_dofs:
                eg = i["element_offset"] + np.arange(lm.t.shape[1])
                l2g[ldn.i_dof] = offset + eg[None, :] * elem.i_dofs + np.arange(elem.i_dofs)[:, None]
            offset += elem.i_dofs * i["n_total"]
            maps.append((l2g, offset))
        self._maps = maps
        return maps

    l2g_u = property(lambda self: self._dof_maps()[0][0])
    l2g_v = property(lambda self: self._dof_maps()[1][0])
    N_u = property(lambda self: self._dof_maps()[0][1])
    N_v = property(lambda self: self._dof_maps()[1][1])


def _match_tuples(local, glob, nv):
    """Index in ``glob`` (k x ng, lexicographically sorted unique columns) of
    every column of ``local`` (k x nl)."""
    k = local.shape[0]
    if float(nv) ** k >= 2.0 ** 62:
        raise NotImplementedError("vertex tuples do not fit a 64-bit key")
    kg = glob[0].astype(np.int64)
    kl = local[0].astype(np.int64)
    for r in range(1, k):
        kg = kg * nv + glob[r]
        kl = kl * nv + local[r]
    pos = np.searchsorted(kg, kl)
    if not np.array_equal(kg[np.minimum(pos, len(kg) - 1)], kl):
        raise ValueError("local entity missing from the global mesh")
    return pos


def slab_part(kind, size, rank, world, elem_u, elem_v=None, halo=True):
    """Rank ``rank``'s part of the weak-scaling meshes of BASELINE configs 3-5:
    ``kind="tri"``: the ``size x (size*world)`` triangle grid of the rectangle
    ``[0,1] x [0,world]`` (``meshgen.grid_tri``), ``kind="tet"``: the Kuhn grid of
    ``[0,1]^2 x [0,world]`` (``meshgen.grid_tet``), cut into ``world`` slabs of
    ``size`` cell layers along the last axis.  Every rank builds only its own
    slab plus (``halo=True``) the layer of cells above it, whose elements touch
    the rows on the slab's top plane -- the global mesh never exists.

    Ownership is the general rule (lowest rank touching the row): a rank owns
    the DOFs of its own cells except those on its bottom plane (rank - 1 owns
    them); DOFs that only the halo layer touches belong to rank + 1.  Works for
    any element (the masks come from the local ``Dofnum``)."""
    from . import meshgen
    k0, k1 = rank * size, (rank + 1) * size
    top = k1 + (1 if (halo and rank < world - 1) else 0)
    if kind == "tri":
        p, t = meshgen.grid_tri(size, ny=size * world, j0=k0, j1=top)
        mesh = _mesh.MeshTri(p, t, validate=False)
        plane = size + 1
        per_layer, axis = 2 * size, 1
        nv_global = (size + 1) * (size * world + 1)
    elif kind == "tet":
        p, t = meshgen.grid_tet(size, mz=size * world, k0=k0, k1=top)
        mesh = _mesh.MeshTet(p, t, validate=False)
        plane = (size + 1) * (size + 1)
        per_layer, axis = 6 * size * size, 2
        nv_global = plane * (size * world + 1)
    else:
        raise ValueError("kind must be 'tri' or 'tet'")
    n_own = per_layer * size
    nv_loc = mesh.p.shape[1]
    v_bottom = np.zeros(nv_loc, dtype=bool)
    if rank > 0:
        v_bottom[:plane] = True                      # first vertex plane of the slab
    parts = []
    for elem in (elem_u, elem_v if (elem_v is not None and elem_v is not elem_u) else None):
        if elem is None:
            parts.append(None)
            continue
        dn = Dofnum(mesh, elem)
        touched = np.zeros(dn.N, dtype=bool)
        touched[dn.t_dof[:, :n_own]] = True          # DOFs of my own cells
        bottom = np.zeros(dn.N, dtype=bool)
        if rank > 0:
            if elem.n_dofs:
                bottom[dn.n_dof[:, v_bottom]] = True
            if hasattr(mesh, "edges") and elem.e_dofs:
                bottom[dn.e_dof[:, v_bottom[mesh.edges].all(axis=0)]] = True
            if elem.f_dofs:
                bottom[dn.f_dof[:, v_bottom[mesh.facets].all(axis=0)]] = True
        owned = touched & ~bottom
        owner = np.where(owned, rank, np.where(bottom, rank - 1, rank + 1)).astype(np.int64)
        parts.append((owned, owner))
    owned_v, owner_v = parts[1] if parts[1] is not None else parts[0]
    info = {"kind": kind, "size": size, "world": world, "rank": rank,
            "vertex_offset": k0 * plane, "nv_global": nv_global,
            "element_offset": k0 * per_layer, "n_total": per_layer * size * world}
    part = SlabPart(mesh, elem_u, elem_v if elem_v is not None else elem_u, owned_v, owner_v,
                    n_own, info)
    hi = float(world)

    def on_global_boundary(lmesh, lf):
        """A local boundary facet is on the global boundary iff its midpoint lies
        on a face of the box [0,1]^(d-1) x [0,world] (the rest are cuts)."""
        mid = lmesh.p[:, lmesh.facets[:, lf]].mean(axis=1)
        on = np.zeros(mid.shape[1], dtype=bool)
        for d in range(mid.shape[0]):
            top_d = hi if d == axis else 1.0
            on |= (mid[d] == 0.0) | (mid[d] == top_d)
        return on

    part.on_global_boundary = on_global_boundary
    return part


def slab_global(kind, size, world):
    """The whole slab mesh (tests and the benchmark's parity check; small sizes)."""
    from . import meshgen
    if kind == "tri":
        p, t = meshgen.grid_tri(size, ny=size * world)
        return _mesh.MeshTri(p, t, validate=False)
    p, t = meshgen.grid_tet(size, mz=size * world)
    return _mesh.MeshTet(p, t, validate=False)


def strip_global(world, refine):
    """The whole strip mesh with the global numbering of :func:`strip_part`
    (tests only; small sizes)."""
    parts = [strip_part(r, world, refine) for r in range(world)]
    N = parts[0].N_v
    P = np.zeros((2, N))
    T = []
    for part in parts:
        P[:, part.l2g_v] = part.mesh.p
        n_own = part.n_own_elements
        # own elements come first in the local mesh (halo appended)
        T.append(part.l2g_v[part.mesh.t[:, :n_own]])
    return _mesh.MeshTri(P, np.hstack(T), validate=False)


class DistResult(object):
    """Row-distributed assembly result.  ``local`` is the rank's local result
    (``engine.DeviceCSR`` / device tensor, or host objects in the CPU tests);
    only the rows flagged in ``part.owned_v`` are meaningful."""

    def __init__(self, local, part, bilinear, group):
        self.local, self.part, self.bilinear, self.group = local, part, bilinear, group

    def owned_triplets(self):
        """Host COO triplets (global ids) of the owned rows."""
        part = self.part
        A = self.local.to_scipy() if hasattr(self.local, "to_scipy") else self.local
        if self.bilinear:
            A = A.tocsr()
            rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
            keep = part.owned_v[rows]
            return (part.l2g_v[rows[keep]], part.l2g_u[A.indices[keep]], A.data[keep])
        v = A.cpu().numpy() if hasattr(A, "cpu") else np.asarray(A)
        idx = np.nonzero(part.owned_v)[0]
        return (part.l2g_v[idx], np.zeros(len(idx), dtype=np.int64), v[idx])

    def csr(self):
        """The rank's row slice on its device: (owned global row ids, ascending;
        ``indptr`` over them; global column ids, sorted inside a row; values) --
        the same layout as :meth:`ExchangedResult.csr`.  Bilinear forms only."""
        import torch
        if not self.bilinear:
            raise NotImplementedError("csr() of a load vector")
        part = self.part
        rows, cols, values = _local_tensors(self.local, True)
        dev = values.device
        keep = torch.from_numpy(np.ascontiguousarray(part.owned_v)).to(dev)[rows]
        l2g_v = torch.from_numpy(np.ascontiguousarray(part.l2g_v, dtype=np.int64)).to(dev)
        l2g_u = torch.from_numpy(np.ascontiguousarray(part.l2g_u, dtype=np.int64)).to(dev)
        ncols = max(int(part.N_u), 1)
        key, order = torch.sort(l2g_v[rows[keep]] * ncols + l2g_u[cols[keep]])
        grow = torch.div(key, ncols, rounding_mode="floor")
        urows, counts = torch.unique_consecutive(grow, return_counts=True)
        indptr = torch.zeros(urows.numel() + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(counts, 0)
        return urows, indptr, key - grow * ncols, values[keep][order]

    def gather(self, dst=0):
        """Collect the global matrix / vector on rank ``dst`` (``None`` on the
        other ranks).  Every row is owned by exactly one rank, so no value is
        ever added across ranks."""
        import torch.distributed as dist
        trip = self.owned_triplets()
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            parts = [trip]
        else:
            rank = dist.get_rank(self.group)
            parts = [None] * world if rank == dst else None
            dist.gather_object(trip, parts, dst=dst, group=self.group)
            if rank != dst:
                return None
        r = np.concatenate([x[0] for x in parts])
        c = np.concatenate([x[1] for x in parts])
        v = np.concatenate([x[2] for x in parts])
        part = self.part
        if self.bilinear:
            return coo_matrix((v, (r, c)), shape=(part.N_v, part.N_u)).tocsr()
        out = np.zeros(part.N_v)
        out[r] = v
        return out


def _local_tensors(loc, bilinear):
    """(row ids, column ids, values) of a rank-local result as torch tensors on
    the device the values live on (``engine.DeviceCSR`` / device vector on the
    GPU, SciPy / NumPy objects in the CPU tests)."""
    import torch
    if bilinear:
        if hasattr(loc, "pattern"):                       # engine.DeviceCSR
            indptr, indices, values = loc.pattern.indptr, loc.pattern.indices, loc.values
        else:
            A = loc.tocsr()
            indptr = torch.from_numpy(A.indptr.astype(np.int64))
            indices = torch.from_numpy(A.indices.astype(np.int64))
            values = torch.from_numpy(np.ascontiguousarray(A.data, dtype=np.float64))
        indptr = indptr.to(torch.int64)
        counts = indptr[1:] - indptr[:-1]
        rows = torch.repeat_interleave(
            torch.arange(counts.numel(), device=values.device, dtype=torch.int64), counts)
        return rows, indices.to(torch.int64), values
    v = loc if isinstance(loc, torch.Tensor) else torch.from_numpy(
        np.ascontiguousarray(loc, dtype=np.float64))
    rows = torch.arange(v.numel(), device=v.device, dtype=torch.int64)
    return rows, None, v                  # (a vector: one global column, id 0)


class InterfaceExchange(object):
    """The north star's exchange step for a ``halo=False`` partition: every rank
    assembles its own element block only; the partial sums of rows it touches
    but does not own ("ghost rows") travel to the owning rank in ONE
    ``all_to_all_single`` (NCCL over NVLink on GPUs, gloo in the CPU tests) and
    the owner adds them **in fixed rank order** -- its own values first, then
    the senders in ascending rank -- so the result is reproducible run to run.

    Built once per (partition, local pattern): the sorted (global row, global
    column) keys of the ghost entries are exchanged, the owner forms the union
    pattern of its rows (an interface row also has columns that only the
    neighbour's elements produce) and every later call moves fp64 values only:
    one gather into the send buffer, one collective, one indexed add per
    sending rank (the indices inside one add are unique: no atomics race).
    Interface volume is O(surface) of the block, so the step is latency bound;
    it is not fused into the element kernel."""

    def __init__(self, part, rows, cols, group=None, rank=None, world=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        dev = rows.device
        self.device = dev
        ncols = max(int(part.N_u), 1) if cols is not None else 1
        self.ncols = ncols
        l2g_v = torch.from_numpy(np.ascontiguousarray(part.l2g_v, dtype=np.int64)).to(dev)
        l2g_u = torch.from_numpy(np.ascontiguousarray(part.l2g_u, dtype=np.int64)).to(dev)
        owner = torch.from_numpy(np.ascontiguousarray(part.owner_v, dtype=np.int64)).to(dev)
        key = l2g_v[rows] * ncols
        if cols is not None:
            key = key + l2g_u[cols]
        own = owner[rows]
        # ---- my ghost entries, grouped by destination, sorted by key inside
        ghost = torch.nonzero(own != self.rank).reshape(-1)
        order = torch.argsort(own[ghost] * (int(part.N_v) * ncols + 1) + key[ghost])
        self.send_idx = ghost[order]
        send_keys = key[self.send_idx].contiguous()
        send_counts = torch.bincount(own[self.send_idx], minlength=self.world)[:self.world]
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        self.send_splits = [int(x) for x in send_counts.tolist()]
        self.recv_splits = [int(x) for x in recv_counts.tolist()]
        recv_keys = torch.empty(sum(self.recv_splits), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv_keys, send_keys, self.recv_splits, self.send_splits,
                               group=group)
        # ---- union pattern of the owned rows
        self.own_idx = torch.nonzero(own == self.rank).reshape(-1)
        own_keys = key[self.own_idx]
        self.keys = torch.unique(torch.cat((own_keys, recv_keys)))      # sorted
        self.pos_own = torch.searchsorted(self.keys, own_keys)
        self.pos_recv = torch.searchsorted(self.keys, recv_keys)
        self.nnz = int(self.keys.numel())
        self.rows = torch.div(self.keys, ncols, rounding_mode="floor")
        self.cols = self.keys - self.rows * ncols
        self.interface_entries = int(self.send_idx.numel())

    def reduce(self, values):
        """Owned-row values (aligned with ``self.rows/self.cols``) from the
        rank-local ``values`` (aligned with the local pattern)."""
        import torch
        import torch.distributed as dist
        send = values.index_select(0, self.send_idx)
        recv = torch.empty(sum(self.recv_splits), dtype=values.dtype, device=values.device)
        dist.all_to_all_single(recv, send, self.recv_splits, self.send_splits,
                               group=self.group)
        out = torch.zeros(self.nnz, dtype=values.dtype, device=values.device)
        out.index_copy_(0, self.pos_own, values.index_select(0, self.own_idx))
        lo = 0
        for s in range(self.world):                       # fixed rank order
            n = self.recv_splits[s]
            if n:
                out.index_add_(0, self.pos_recv[lo:lo + n], recv[lo:lo + n])
            lo += n
        return out


class ExchangedResult(object):
    """Row-distributed result after the interface exchange: the owned rows with
    *global* row / column ids (sorted by row, then column) and their values on
    the rank's device.  ``csr()`` gives the rank's row slice as device CSR
    arrays; ``gather`` collects the global host object like
    :meth:`DistResult.gather`."""

    def __init__(self, xchg, values, part, bilinear, group):
        self.xchg, self.values, self.part = xchg, values, part
        self.bilinear, self.group = bilinear, group

    def csr(self):
        """(owned global row ids, indptr over them, global column ids, values)."""
        import torch
        urows, counts = torch.unique_consecutive(self.xchg.rows, return_counts=True)
        indptr = torch.zeros(urows.numel() + 1, dtype=torch.int64, device=urows.device)
        indptr[1:] = torch.cumsum(counts, 0)
        return urows, indptr, self.xchg.cols, self.values

    def owned_triplets(self):
        return (self.xchg.rows.cpu().numpy(), self.xchg.cols.cpu().numpy(),
                self.values.cpu().numpy())

    gather = DistResult.gather


class DistributedAssembler(object):
    """``AssemblerElement`` over an element-partitioned mesh, one rank per GPU.

    ``DistributedAssembler(mesh, elem_u, elem_v=None)`` partitions the
    replicated host ``mesh`` (every rank passes the same object) for this
    process' rank; ``part=`` supplies a ready :class:`LocalPart` instead
    (e.g. :func:`strip_part`).  ``iasm`` / ``fasm`` have the reference's
    signatures and return a :class:`DistResult`."""

    def __init__(self, mesh, elem_u, elem_v=None, group=None, part=None,
                 rank=None, world=None, assemble=None, mode="halo",
                 assemble_facet=None, facet_halo=False):
        import torch.distributed as dist
        if mode not in ("halo", "exchange"):
            raise ValueError("mode must be 'halo' or 'exchange'")
        self.mode = mode
        self._xchg = {}
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world, self.group = rank, world, group
        self.facet_halo = facet_halo
        self.part = part if part is not None else partition_global(
            mesh, elem_u, elem_v, world, rank, halo=(mode == "halo"),
            facet_halo=facet_halo)
        if mode == "exchange" and self.part.owner_v is None:
            raise ValueError("mode='exchange' needs a partition with owner_v "
                             "(partition_global(..., halo=False))")
        self.local = AssemblerElement(self.part.mesh, elem_u, elem_v)
        if mode == "halo":
            # rows another rank owns are not assembled here at all (the native tile
            # plan gives them no slots; the other routes compute and ignore them)
            self.local.row_mask = np.ascontiguousarray(self.part.owned_v, dtype=bool)
        self._assemble = assemble
        self._assemble_facet = assemble_facet

    def iasm(self, form, intorder=None, interp=None, to_host=False):
        if interp is not None:
            interp = {k: np.asarray(v)[self.part.l2g_u] for k, v in interp.items()}
        bilinear = self.local.plan_iasm(form, intorder=intorder, interp=interp).bilinear
        if self._assemble is not None:          # CPU tests inject the local assembly
            loc = self._assemble(self.local, form, intorder=intorder, interp=interp)
        elif to_host:
            loc = self.local.iasm(form, intorder=intorder, interp=interp)
        else:
            loc = self.local.iasm_device(form, intorder=intorder, interp=interp)
        if self.mode == "exchange":
            return self._exchange(loc, bilinear)
        return DistResult(loc, self.part, bilinear, self.group)

    def _exchange(self, loc, bilinear, kind="interior"):
        """Send the ghost rows' partial sums to their owners (one collective)."""
        rows, cols, values = _local_tensors(loc, bilinear)
        # one exchange plan per local pattern (the pattern object is cached by
        # the engine per (row table, column table); host results: per shape/nnz)
        pkey = (kind, bilinear, id(getattr(loc, "pattern", None)), int(values.numel()))
        hit = self._xchg.get(pkey)
        if hit is None:                # (the pattern is kept alive with its plan)
            while len(self._xchg) >= 8:            # bounded: oldest plan out
                self._xchg.pop(next(iter(self._xchg)))
            hit = self._xchg[pkey] = (InterfaceExchange(
                self.part, rows, cols, self.group, self.rank, self.world),
                getattr(loc, "pattern", None))
        x = hit[0]
        return ExchangedResult(x, x.reduce(values), self.part, bilinear, self.group)

    def fasm(self, form, find=None, interior=False, intorder=None, normals=True,
             interp=None, to_host=False):
        """Boundary-facet assembly over the partitioned mesh (the Nitsche terms
        of BASELINE config 5).  ``find`` is not supported (facet ids are local
        to a rank): the facets are the rank's share of the *global* boundary --
        the artificial boundary along the cuts is left out.  Halo mode: a rank
        assembles the boundary facets of its own and its halo elements, which
        completes its owned rows; exchange mode: of its own elements only, then
        the ghost rows travel to their owners.

        ``interior=True`` (two-sided DG / jump forms): needs the element across
        every facet of an element that touches an owned row, i.e. a partition
        built with ``facet_halo=True`` (halo mode); the rank then assembles all
        interior facets of its local mesh -- the cut facets are boundary facets
        of the local mesh and belong to the neighbour that has both sides."""
        if interior and not (self.mode == "halo" and self.facet_halo):
            raise NotImplementedError(
                "interior facets over a partitioned mesh need "
                "DistributedAssembler(..., mode='halo', facet_halo=True)")
        if find is not None:
            raise NotImplementedError("find= over a partitioned mesh (facet ids are rank-local)")
        if interp is not None:
            interp = {k: np.asarray(v)[self.part.l2g_u] for k, v in interp.items()}
        lf = (self.part.mesh.interior_facets() if interior
              else self.part.global_boundary_facets())
        plan = self.local.plan_fasm(form, interior=interior, intorder=intorder,
                                    normals=normals, interp=interp)
        if self._assemble_facet is not None:    # CPU tests inject the local assembly
            loc = self._assemble_facet(self.local, form, find=lf, interior=interior,
                                       intorder=intorder, normals=normals, interp=interp)
        else:
            loc = self.local._get_engine().run_facet(plan, np.asarray(lf), interior, interp,
                                                     to_host=to_host)
        if self.mode == "exchange":
            return self._exchange(loc, plan.bilinear, kind="facet")
        return DistResult(loc, self.part, plan.bilinear, self.group)

    def prepare_iasm(self, form, intorder=None):
        """Staged local assembly (see ``AssemblerElement.prepare_iasm``)."""
        return self.local.prepare_iasm(form, intorder=intorder)
