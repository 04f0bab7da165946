# -*- coding: utf-8 -*-
"""Host helpers kept from spfem/utils.py: nested-dict "cell arrays" used by
facet forms (``const_cell`` :30-50, ``cell_shape`` :22-28), block stacking
(:12-20) and the Dirichlet-eliminating direct solve (:52-95) that the README
example ends with.  Solvers are consumers of the hot path, not part of it."""
from copy import deepcopy
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl


def stack(block):
    """Stack a 2-D tuple array of scipy sparse matrices."""
    return sp.vstack([sp.hstack(row) for row in block])


def cell_shape(x, *rest):
    """Shape of a nested dict ("cell array")."""
    if isinstance(x, dict):
        return cell_shape(x[0], len(x), *rest)
    return rest[::-1]


def const_cell(nparr, *arg):
    """Nested dict of the given shape filled with deep copies of ``nparr``."""
    if len(arg) == 0:
        return nparr
    if len(arg) == 1:
        return {i: deepcopy(nparr) for i in range(arg[0])}
    return {i: const_cell(nparr, *arg[1:]) for i in range(arg[0])}


def direct(A, b, x=None, I=None, use_umfpack=True, cholmod=False):
    """Solve ``Ax = b``, optionally only on the index set ``I`` with the
    remaining entries of ``x`` as Dirichlet data (spfem/utils.py:52-95)."""
    if cholmod:
        raise NotImplementedError("cholmod (scikit-sparse) is not available")
    if I is None:
        return spl.spsolve(A, b, use_umfpack=use_umfpack)
    AII = A[I].T[I].T
    if x is None:
        x = np.zeros(A.shape[0])
        x[I] = spl.spsolve(AII, b[I], use_umfpack=use_umfpack)
    else:
        D = np.setdiff1d(np.arange(A.shape[0]), I)
        x[I] = spl.spsolve(AII, b[I] - A[I].T[D].T.dot(x[D]),
                           use_umfpack=use_umfpack)
    return x


def gradient(u, mesh):
    """Elementwise gradient of a P1 function on a triangular mesh."""
    x1, x2, x3 = (mesh.p[0, mesh.t[k, :]] for k in range(3))
    y1, y2, y3 = (mesh.p[1, mesh.t[k, :]] for k in range(3))
    z1, z2, z3 = (u[mesh.t[k, :]] for k in range(3))
    det = x2*y1 - x3*y1 - x1*y2 + x3*y2 + x1*y3 - x2*y3
    dx = (-y2*z1 + y3*z1 + y1*z2 - y3*z2 - y1*z3 + y2*z3) / det
    dy = (x2*z1 - x3*z1 - x1*z2 + x3*z2 + x1*z3 - x2*z3) / det
    return dx, dy
