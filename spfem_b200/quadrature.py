# -*- coding: utf-8 -*-
"""Quadrature rules on the reference domains (drop-in for spfem/quadrature.py).

``get_quadrature(refdom, norder)`` mirrors spfem/quadrature.py:7-62: tabulated
triangle (orders 2..19) and tetrahedron (orders 2..4) rules, Gauss-Legendre on
[0, 1] for lines and its tensor product mapped to [-1, 1]^2 for quads.  The
tabulated numbers live in ``data/quadrature.json`` (regenerated from the
reference by ``tools/gen_quadrature_tables.py``); the same arrays are what the
generated kernels embed as compile-time constants.
"""
import json
import os
import numpy as np

_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                            "data", "quadrature.json")
        with open(path) as fh:
            raw = json.load(fh)
        _TABLES = {dom: {int(k): (np.array(v["X"], dtype=np.float64),
                                  np.array(v["W"], dtype=np.float64))
                         for k, v in rules.items()}
                   for dom, rules in raw.items()}
    return _TABLES


def get_quadrature_tri(norder):
    """Reference triangle (0,0) (1,0) (0,1); spfem/quadrature.py:87-114."""
    if norder <= 1:
        norder = 2
    try:
        X, W = _tables()["tri"][int(norder)]
    except (KeyError, ValueError, TypeError):
        raise NotImplementedError("get_quadrature_tri: The requested order "
                                  "of quadrature is not implemented!")
    return X.copy(), W.copy()


def get_quadrature_tet(norder):
    """Reference tetrahedron; spfem/quadrature.py:64-86 (orders 2..4 only)."""
    if norder <= 1:
        norder = 2
    try:
        X, W = _tables()["tet"][int(norder)]
    except (KeyError, ValueError, TypeError):
        raise NotImplementedError("The requested order of quadrature "
                                  "is not tabulated.")
    return X.copy(), W.copy()


def get_quadrature_line(norder):
    """Gauss-Legendre on [0, 1], weights sum to 1; spfem/quadrature.py:116-121."""
    if norder <= 1:
        norder = 2
    npts = int(np.ceil((norder + 1.0) / 2.0))
    X, W = np.polynomial.legendre.leggauss(npts)
    return np.array([0.5 * X + 0.5]), W / 2.0


def get_quadrature(refdom, norder):
    """Return ``(X, W)``: points ``dim x nqp`` and weights ``nqp``."""
    if refdom == "tri":
        return get_quadrature_tri(norder)
    elif refdom == "tet":
        return get_quadrature_tet(norder)
    elif refdom == "line":
        return get_quadrature_line(norder)
    elif refdom == "quad":
        # tensor rule on (-1,-1) (1,-1) (1,1) (-1,1); point order is the
        # column-major flattening of meshgrid (spfem/quadrature.py:51-60)
        X, W = get_quadrature_line(norder)
        A, B = np.meshgrid(X, X)
        Y = 2.0 * np.vstack((A.flatten(order='F'), B.flatten(order='F'))) - 1.0
        A, B = np.meshgrid(2 * W, 2 * W)
        return Y, (A * B).flatten(order='F')
    raise NotImplementedError("The given mesh type is not supported!")
