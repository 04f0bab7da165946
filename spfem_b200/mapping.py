# -*- coding: utf-8 -*-
"""Reference <-> global maps (drop-in for spfem/mapping.py).

``MappingAffine`` (simplices; spfem/mapping.py:193-567) and ``MappingQ1``
(quadrilaterals; :47-191) keep the reference's public surface -- ``A, b, detA,
invA, B, c, detB`` and ``F, invF, G, DF, invDF, detDF, detDG, normals`` -- as
*host* helpers for user code and error norms.  The assembly kernels do not read
these arrays: they recompute the same quantities on the device from ``p`` and
``t`` with the same operation order (see ``codegen.py``), which is why the
per-element host arrays here are built lazily on first access.
"""
import numpy as np


def _rows(X, dim):
    """Reference points as a list of ``dim`` arrays, from a (dim x n) array or
    a dict of (n_elems x nqp) arrays."""
    return [X[k] for k in range(dim)]


def _npts(X):
    return X[0].shape[1] if isinstance(X, dict) else X.shape[1]


class Mapping(object):
    """Abstract mapping (spfem/mapping.py:14-45)."""
    dim = 0

    def _count(self, tind):
        return self._nt if tind is None else len(tind)


class MappingAffine(Mapping):
    """``F(X) = A X + b`` per element, ``G(X) = B X + c`` per facet."""

    _LAZY = ("A", "b", "detA", "invA", "B", "c", "detB")

    def __init__(self, mesh):
        from . import mesh as fmsh
        if isinstance(mesh, fmsh.MeshLine):
            self.dim = 1
        elif isinstance(mesh, fmsh.MeshTri):
            self.dim = 2
        elif isinstance(mesh, fmsh.MeshTet):
            self.dim = 3
        else:
            raise TypeError("MappingAffine initialized with an incompatible "
                            "mesh type!")
        self._mesh = mesh
        self._nt = mesh.t.shape[1]

    def __getattr__(self, name):
        if name in MappingAffine._LAZY:
            self._compute()
            return self.__dict__[name]
        raise AttributeError(name)

    def _compute(self):
        """Per-element / per-facet geometry (spfem/mapping.py:195-305)."""
        p, t, d = self._mesh.p, self._mesh.t, self.dim
        if d == 1:
            self.A = p[0, t[1, :]] - p[0, t[0, :]]
            self.b = p[0, t[0, :]]
            self.detA = self.A
            self.invA = 1.0 / self.A
            self.B = self.c = self.detB = None
            return
        A = {i: {j: p[i, t[j + 1, :]] - p[i, t[0, :]] for j in range(d)}
             for i in range(d)}
        self.A = A
        self.b = {i: p[i, t[0, :]] for i in range(d)}
        fa = self._mesh.facets
        if d == 2:
            det = A[0][0] * A[1][1] - A[0][1] * A[1][0]
            self.invA = {0: {0: A[1][1] / det, 1: -A[0][1] / det},
                         1: {0: -A[1][0] / det, 1: A[0][0] / det}}
            self.B = {i: p[i, fa[1, :]] - p[i, fa[0, :]] for i in range(2)}
            self.detB = np.sqrt(self.B[0] ** 2 + self.B[1] ** 2)
        else:
            det = (A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1])
                   - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0])
                   + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]))
            inv = {0: {}, 1: {}, 2: {}}
            inv[0][0] = (-A[1][2] * A[2][1] + A[1][1] * A[2][2]) / det
            inv[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det
            inv[2][0] = (-A[1][1] * A[2][0] + A[1][0] * A[2][1]) / det
            inv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det
            inv[1][1] = (-A[0][2] * A[2][0] + A[0][0] * A[2][2]) / det
            inv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det
            inv[0][2] = (-A[0][2] * A[1][1] + A[0][1] * A[1][2]) / det
            inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det
            inv[2][2] = (-A[0][1] * A[1][0] + A[0][0] * A[1][1]) / det
            self.invA = inv
            B = {i: {j: p[i, fa[j + 1, :]] - p[i, fa[0, :]] for j in range(2)}
                 for i in range(3)}
            self.B = B
            cr0 = B[1][0] * B[2][1] - B[2][0] * B[1][1]
            cr1 = -B[0][0] * B[2][1] + B[2][0] * B[0][1]
            cr2 = B[0][0] * B[1][1] - B[1][0] * B[0][1]
            self.detB = np.sqrt(cr0 ** 2 + cr1 ** 2 + cr2 ** 2)
        self.detA = det
        self.c = {i: p[i, fa[0, :]] for i in range(d)}

    @staticmethod
    def _sel(arr, ind):
        return arr if ind is None else arr[ind]

    def F(self, X, tind=None):
        """Reference -> global, (n_elems x nqp) per coordinate."""
        d = self.dim
        if d == 1:
            return (np.outer(self.A, X[0, :]).T + self.b).T
        y = {}
        for i in range(d):
            acc = np.outer(self._sel(self.A[i][0], tind), X[0, :]).T
            for j in range(1, d):
                acc = acc + np.outer(self._sel(self.A[i][j], tind), X[j, :]).T
            y[i] = (acc + self._sel(self.b[i], tind)).T
        return y

    def invF(self, x, tind=None):
        """Global -> reference, ``A^{-1}(x - b)``."""
        d = self.dim
        if d not in (2, 3):
            raise NotImplementedError("MappingAffine.F: given dimension not "
                                      "implemented yet!")
        Y = {i: x[i].T - self._sel(self.b[i], tind) for i in range(d)}
        y = {}
        for i in range(d):
            acc = self._sel(self.invA[i][0], tind) * Y[0]
            for j in range(1, d):
                acc = acc + self._sel(self.invA[i][j], tind) * Y[j]
            y[i] = acc.T
        return y

    def G(self, X, find=None):
        """Reference facet -> global facet."""
        y = {}
        if self.dim == 2:
            for i in range(2):
                y[i] = (np.outer(self._sel(self.B[i], find), X).T
                        + self._sel(self.c[i], find)).T
        elif self.dim == 3:
            for i in range(3):
                y[i] = (np.outer(self._sel(self.B[i][0], find), X[0, :]).T
                        + np.outer(self._sel(self.B[i][1], find), X[1, :]).T
                        + self._sel(self.c[i], find)).T
        else:
            raise NotImplementedError("MappingAffine.G: given dimension not "
                                      "implemented yet!")
        return y

    def _tiled(self, table, X, ind):
        N = _npts(X)
        d = self.dim
        return {i: {j: np.tile(self._sel(table[i][j], ind), (N, 1)).T
                    for j in range(d)} for i in range(d)}

    def DF(self, X, tind=None):
        return self._tiled(self.A, X, tind)

    def invDF(self, X, tind=None):
        if self.dim == 1:
            return self._sel(self.invA, tind)
        return self._tiled(self.invA, X, tind)

    def detDF(self, X, tind=None):
        return np.tile(self._sel(self.detA, tind), (_npts(X), 1)).T

    def detDG(self, X, find=None):
        return np.tile(self._sel(self.detB, find), (X.shape[1], 1)).T

    def normals(self, X, tind, find, t2f):
        """Unit outward normals of element ``tind`` on facets ``find``:
        ``normalize(DF^{-T} n_ref)`` (spfem/mapping.py:474-524)."""
        if self.dim == 2:
            nref = np.array([[0.0, -1.0], [1.0, 1.0], [-1.0, 0.0]])
        elif self.dim == 3:
            nref = np.array([[0.0, 0.0, -1.0], [0.0, -1.0, 0.0],
                             [-1.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
        else:
            raise NotImplementedError("MappingAffine.normals() not implemented "
                                      "for the used self.dim.")
        d = self.dim
        invDF = self.invDF(X, tind)
        n = {k: np.zeros(find.shape[0]) for k in range(d)}
        for itr in range(nref.shape[0]):
            inds = np.nonzero(t2f[itr, tind] == find)[0]
            for jtr in range(d):
                n[jtr][inds] = nref[itr, jtr]
        N = {}
        for k in range(d):
            acc = invDF[0][k] * n[0][:, None]
            for l in range(1, d):
                acc = acc + invDF[l][k] * n[l][:, None]
            N[k] = acc
        nlen = N[0] ** 2 + N[1] ** 2
        if d == 3:
            nlen = nlen + N[2] ** 2
        nlen = np.sqrt(nlen)
        return {k: N[k] / nlen for k in range(d)}


class MappingQ1(Mapping):
    """Bilinear map of [-1,1]^2 onto each quadrilateral
    (spfem/mapping.py:47-191)."""

    def __init__(self, mesh):
        from . import mesh as fmsh
        if not isinstance(mesh, fmsh.MeshQuad):
            raise NotImplementedError("MappingQ1: wrong type of mesh was "
                                      "given to constructor!")
        self.dim = 2
        self.t = mesh.t
        self.p = mesh.p
        self._nt = mesh.t.shape[1]
        fa = mesh.facets
        self.B = {i: mesh.p[i, fa[1, :]] - mesh.p[i, fa[0, :]] for i in range(2)}
        self.c = {i: mesh.p[i, fa[0, :]] for i in range(2)}
        self.detB = np.sqrt(self.B[0] ** 2 + self.B[1] ** 2)
        # dF_i/dX_j as closures (x, y, tind), kept for API parity (:58-74)
        self.J = {i: {j: (lambda x, y, t, i=i, j=j: self._jac(i, j, x, y, t))
                      for j in range(2)} for i in range(2)}

    def _corner(self, i, k, tind):
        return self.p[i, self.t[k, tind]][:, None]

    def _jac(self, i, j, x, y, tind):
        P = [self._corner(i, k, tind) for k in range(4)]
        if j == 0:
            return 0.25 * (-P[0] * (1 - y) + P[1] * (1 - y)
                           + P[2] * (1 + y) - P[3] * (1 + y))
        return 0.25 * (-P[0] * (1 - x) - P[1] * (1 + x)
                       + P[2] * (1 + x) + P[3] * (1 - x))

    @staticmethod
    def quadbasis(x, y, i):
        sx = (-1, 1, 1, -1)[i]
        sy = (-1, -1, 1, 1)[i]
        return 0.25 * (1 + sx * x) * (1 + sy * y)

    def _tind(self, tind):
        return np.arange(self.t.shape[1]) if tind is None else tind

    def F(self, Y, tind=None):
        X = Y if isinstance(Y, dict) else {0: Y[0, :], 1: Y[1, :]}
        tind = self._tind(tind)
        out = {}
        for i in range(2):
            acc = self._corner(i, 0, tind) * self.quadbasis(X[0], X[1], 0)
            for k in range(1, 4):
                acc = acc + self._corner(i, k, tind) * self.quadbasis(X[0], X[1], k)
            out[i] = acc
        return out

    def invF(self, x, tind=None):
        """One Newton step from the origin (spfem/mapping.py:124-139)."""
        X = {0: 0 * x[0], 1: 0 * x[1]}
        F = self.F(X, tind)
        invDF = self.invDF(X, tind)
        g = {0: x[0] - F[0], 1: x[1] - F[1]}
        return {0: X[0] + invDF[0][0] * g[0] + invDF[0][1] * g[1],
                1: X[1] + invDF[1][0] * g[0] + invDF[1][1] * g[1]}

    def _xy(self, X):
        if isinstance(X, dict):
            return X[0], X[1]
        return X[0, :], X[1, :]

    def detDF(self, X, tind=None):
        tind = self._tind(tind)
        x, y = self._xy(X)
        return (self._jac(0, 0, x, y, tind) * self._jac(1, 1, x, y, tind)
                - self._jac(0, 1, x, y, tind) * self._jac(1, 0, x, y, tind))

    def invDF(self, X, tind=None):
        tind = self._tind(tind)
        x, y = self._xy(X)
        det = self.detDF(X, tind)
        return {0: {0: self._jac(1, 1, x, y, tind) / det,
                    1: -self._jac(0, 1, x, y, tind) / det},
                1: {0: -self._jac(1, 0, x, y, tind) / det,
                    1: self._jac(0, 0, x, y, tind) / det}}

    def G(self, X, find=None):
        sel = MappingAffine._sel
        return {i: (np.outer(sel(self.B[i], find), X).T
                    + sel(self.c[i], find)).T for i in range(2)}

    def detDG(self, X, find=None):
        return np.tile(MappingAffine._sel(self.detB, find), (X.shape[1], 1)).T
