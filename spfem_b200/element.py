# -*- coding: utf-8 -*-
"""Finite elements of the assembly hot path (drop-in for spfem/element.py).

Every element is a list of *polynomials* on its reference domain
(:class:`Poly`, exact rational coefficients).  The same polynomial object is

* evaluated with NumPy for the host API (``lbasis`` / ``gbasis``,
  spfem/element.py:454-489, 491-546),
* tabulated at the quadrature points and embedded as constants in the generated
  sm_100a kernels (interior assembly), and
* emitted as CUDA C and evaluated at per-facet reference points (facet
  assembly, where the points differ from facet to facet).

The nodal bases are *derived* (Vandermonde solve in exact arithmetic on the
element's nodes) rather than typed in; the node/edge ordering is the
reference's: vertices first, then edge/facet midpoints in mesh-local order
(spfem/element.py:879, 1013-1020; spfem/mesh.py:451,820-827), Q2: 4 vertices,
4 edge midpoints (0,1)(1,2)(2,3)(0,3), centre (spfem/element.py:586-596).
"""
from fractions import Fraction
from itertools import product
import numpy as np


class Poly(object):
    """Multivariate polynomial ``sum_e c_e * prod_k x_k**e_k`` with exact
    coefficients; ``terms`` maps exponent tuples to Fractions."""

    def __init__(self, terms, nvar):
        self.nvar = nvar
        self.terms = {tuple(e): Fraction(c) for e, c in terms.items()
                      if Fraction(c) != 0}

    def diff(self, k):
        out = {}
        for e, c in self.terms.items():
            if e[k] > 0:
                ne = list(e)
                ne[k] -= 1
                ne = tuple(ne)
                out[ne] = out.get(ne, 0) + c * e[k]
        return Poly(out, self.nvar)

    def __mul__(self, other):
        out = {}
        for e1, c1 in self.terms.items():
            for e2, c2 in other.terms.items():
                e = tuple(a + b for a, b in zip(e1, e2))
                out[e] = out.get(e, 0) + c1 * c2
        return Poly(out, self.nvar)

    def __add__(self, other):
        out = dict(self.terms)
        for e, c in other.terms.items():
            out[e] = out.get(e, 0) + c
        return Poly(out, self.nvar)

    def __sub__(self, other):
        return self + other.scale(-1)

    def scale(self, c):
        c = Fraction(c)
        return Poly({e: v * c for e, v in self.terms.items()}, self.nvar)

    def is_constant(self):
        return all(sum(e) == 0 for e in self.terms)

    def constant_value(self):
        return float(sum(self.terms.values(), Fraction(0)))

    def sorted_terms(self):
        return sorted(self.terms.items(), key=lambda ec: (sum(ec[0]), ec[0]))

    def __call__(self, *xs):
        """NumPy evaluation; the result always has the shape of ``xs[0]``."""
        out = np.zeros(np.shape(xs[0]), dtype=np.float64)
        for e, c in self.sorted_terms():
            term = float(c)
            for k, ek in enumerate(e):
                for _ in range(ek):
                    term = term * xs[k]
            out = out + term
        return out

    def c_expr(self, names):
        """CUDA C expression in the variables ``names``."""
        if not self.terms:
            return "0.0"
        parts = []
        for e, c in self.sorted_terms():
            fac = [repr(float(c))]
            for k, ek in enumerate(e):
                fac += [names[k]] * ek
            parts.append("*".join(fac))
        return "(" + " + ".join(parts) + ")"


def _solve_exact(V):
    """Inverse of a small square Fraction matrix (Gauss-Jordan)."""
    n = len(V)
    M = [list(row) + [Fraction(int(r == c)) for c in range(n)]
         for r, row in enumerate(V)]
    for col in range(n):
        piv = next(r for r in range(col, n) if M[r][col] != 0)
        M[col], M[piv] = M[piv], M[col]
        pv = M[col][col]
        M[col] = [v / pv for v in M[col]]
        for r in range(n):
            if r != col and M[r][col] != 0:
                f = M[r][col]
                M[r] = [a - f * b for a, b in zip(M[r], M[col])]
    return [row[n:] for row in M]


def nodal_basis(exponents, nodes):
    """Lagrange basis of span{x**e} w.r.t. point evaluation at ``nodes``."""
    nvar = len(nodes[0])
    V = []
    for nd in nodes:
        row = []
        for e in exponents:
            val = Fraction(1)
            for k in range(nvar):
                val *= Fraction(nd[k]) ** e[k]
            row.append(val)
        V.append(row)
    inv = _solve_exact(V)
    return [Poly({e: inv[c][i] for c, e in enumerate(exponents)}, nvar)
            for i in range(len(nodes))]


def _simplex_exponents(dim, deg):
    return [e for e in product(range(deg + 1), repeat=dim) if sum(e) <= deg]


def _simplex_nodes(dim, deg, local_edges):
    H = Fraction(1, 2)
    verts = [tuple(Fraction(int(k == d - 1)) for k in range(dim))
             for d in range(dim + 1)]
    nodes = list(verts)
    if deg == 2:
        for a, b in local_edges:
            nodes.append(tuple(H * (verts[a][k] + verts[b][k])
                               for k in range(dim)))
    return nodes


class Element(object):
    """A finite element defined through basis functions
    (spfem/element.py:14-31)."""

    maxdeg = 0   #: maximum polynomial degree (drives the default quadrature)
    dim = 0      #: spatial dimension
    n_dofs = 0   #: DOFs per vertex
    i_dofs = 0   #: interior DOFs
    f_dofs = 0   #: DOFs per facet
    e_dofs = 0   #: DOFs per edge (3-D only)

    def lbasis(self, X, i):
        raise NotImplementedError("Local basis (lbasis) not implemented!")

    def gbasis(self, mapping, X, i, tind):
        raise NotImplementedError("Global basis (gbasis) not implemented!")


def _coords(X, dim):
    if isinstance(X, dict):
        return [X[k] for k in range(dim)]
    return [X[k] for k in range(dim)]


class ElementH1(Element):
    """H1-conforming element given by reference polynomials ``_polys``; the
    global basis is the push-forward ``grad u = DF^{-T} grad_ref phi``
    (spfem/element.py:457-489)."""

    _polys = ()
    _dpolys = None

    def polys(self):
        """``(phi_i, [d phi_i / dX_k])`` for every local basis function."""
        if self._dpolys is None:
            self._dpolys = [[p.diff(k) for k in range(self.dim)]
                            for p in self._polys]
        return self._polys, self._dpolys

    def nbfun_ref(self):
        return len(self._polys)

    def lbasis(self, X, i):
        i = int(i)
        P, dP = self.polys()
        xs = _coords(X, self.dim)
        phi = P[i](*xs)
        if self.dim == 1:
            return phi, dP[i][0](*xs)
        return phi, {k: dP[i][k](*xs) for k in range(self.dim)}

    def gbasis(self, mapping, X, i, tind):
        dim = mapping.dim
        if isinstance(X, dict):
            phi, dphi = self.lbasis(X, i)
            u = phi
        else:
            x = {k: X[k, :] for k in range(dim)}
            phi, dphi = self.lbasis(x, i)
            n = mapping._count(tind)
            u = np.tile(phi, (n, 1))
        invDF = mapping.invDF(X, tind)
        self.dim = dim
        if dim == 1:
            return u, np.outer(invDF, dphi)
        du = {}
        for k in range(dim):
            acc = invDF[0][k] * dphi[0]
            for l in range(1, dim):
                acc = acc + invDF[l][k] * dphi[l]
            du[k] = acc
        return u, du


class ElementH1Vec(ElementH1):
    """Vector-valued version of a scalar H1 element: local basis ``i`` is
    scalar function ``i // dim`` in component ``i % dim`` (interleaved
    components; spfem/element.py:491-546)."""

    def __init__(self, elem):
        if elem.dim == 0:
            print("ElementH1Vec.__init__(): Warning! Parent element has no dim-variable!")
        self.dim = elem.dim
        self.elem = elem
        self.n_dofs = elem.n_dofs * self.dim
        self.f_dofs = elem.f_dofs * self.dim
        self.i_dofs = elem.i_dofs * self.dim
        self.e_dofs = elem.e_dofs * self.dim
        self.maxdeg = elem.maxdeg

    def polys(self):
        return self.elem.polys()

    def nbfun_ref(self):
        return self.elem.nbfun_ref() * self.dim

    def lbasis(self, X, i):
        return self.elem.lbasis(X, i)

    def gbasis(self, mapping, X, i, tind):
        ind = int(i) // self.dim
        n = int(i) - self.dim * ind
        dim = mapping.dim
        if isinstance(X, dict):
            phi, dphi = self.elem.lbasis(X, ind)
        else:
            x = {k: X[k, :] for k in range(dim)}
            phi, dphi = self.elem.lbasis(x, ind)
            cnt = mapping._count(tind)
            phi = np.tile(phi, (cnt, 1))
            dphi = {k: np.tile(dphi[k], (cnt, 1)) for k in range(dim)}
        u, du = {}, {}
        for c in range(self.dim):
            if c == n:
                u[c] = phi
                invDF = mapping.invDF(X, tind)
                du[c] = {}
                for k in range(dim):
                    acc = invDF[0][k] * dphi[0]
                    for l in range(1, dim):
                        acc = acc + invDF[l][k] * dphi[l]
                    du[c][k] = acc
            else:
                u[c] = 0 * phi
                du[c] = {k: 0 * phi for k in range(self.dim)}
        return u, du


# ---------------------------------------------------------------------------
# concrete elements
# ---------------------------------------------------------------------------
_TRI_EDGES = ((0, 1), (1, 2), (0, 2))
_TET_EDGES = ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))


class ElementTriP1(ElementH1):
    """Linear triangle (spfem/element.py:977-1003)."""
    n_dofs = 1
    dim = 2
    maxdeg = 1
    _polys = nodal_basis(_simplex_exponents(2, 1), _simplex_nodes(2, 1, ()))


class ElementTriP2(ElementH1):
    """Quadratic triangle: 3 vertex + 3 facet DOFs (spfem/element.py:1005-1041)."""
    n_dofs = 1
    f_dofs = 1
    dim = 2
    maxdeg = 2
    _polys = nodal_basis(_simplex_exponents(2, 2),
                         _simplex_nodes(2, 2, _TRI_EDGES))


class ElementTetP1(ElementH1):
    """Linear tetrahedron (spfem/element.py:1043-1079)."""
    n_dofs = 1
    dim = 3
    maxdeg = 1
    _polys = nodal_basis(_simplex_exponents(3, 1), _simplex_nodes(3, 1, ()))


class ElementTetP2(ElementH1):
    """Quadratic tetrahedron: 4 vertex + 6 edge DOFs (spfem/element.py:869-930)."""
    n_dofs = 1
    e_dofs = 1
    dim = 3
    maxdeg = 2
    _polys = nodal_basis(_simplex_exponents(3, 2),
                         _simplex_nodes(3, 2, _TET_EDGES))


_Q_VERTS = ((-1, -1), (1, -1), (1, 1), (-1, 1))


class ElementQ1(ElementH1):
    """Bilinear quadrilateral on [-1,1]^2 (spfem/element.py:548-575);
    ``maxdeg = 2`` as in the reference."""
    maxdeg = 2
    n_dofs = 1
    dim = 2
    _polys = nodal_basis(list(product(range(2), repeat=2)), list(_Q_VERTS))


class ElementQ2(ElementH1):
    """Biquadratic quadrilateral (spfem/element.py:577-621); ``maxdeg = 3``
    as in the reference."""
    maxdeg = 3
    n_dofs = 1
    f_dofs = 1
    i_dofs = 1
    dim = 2
    _polys = nodal_basis(list(product(range(3), repeat=2)),
                         list(_Q_VERTS) + [(0, -1), (1, 0), (0, 1), (-1, 0),
                                           (0, 0)])


class ElementTriP0(ElementH1):
    """Piecewise constants on triangles (spfem/element.py:813-832)."""
    i_dofs = 1
    maxdeg = 1
    dim = 2
    _polys = [Poly({(0, 0): 1}, 2)]


class ElementP0(ElementTriP0):
    pass


class ElementTetP0(ElementH1):
    """Piecewise constants on tetrahedra (spfem/element.py:791-811)."""
    i_dofs = 1
    maxdeg = 1
    dim = 3
    _polys = [Poly({(0, 0, 0): 1}, 3)]


class ElementTriMini(ElementH1):
    """MINI element: P1 + cubic bubble (spfem/element.py:838-867)."""
    dim = 2
    n_dofs = 1
    i_dofs = 1
    maxdeg = 3
    _polys = list(ElementTriP1._polys) + [
        ElementTriP1._polys[0] * ElementTriP1._polys[1] * ElementTriP1._polys[2]]


class ElementHdiv(Element):
    """H(div)-conforming elements (spfem/element.py:404-425): the reference
    basis ``(phi_hat, div phi_hat)`` is pushed forward with the contravariant
    Piola map ``u = DF phi_hat / detDF``, ``div u = div phi_hat / detDF``
    (signed determinant).  ``_vec[i]`` are the component polynomials and
    ``_div[i]`` the divergence of local basis function ``i``."""
    hdiv = True
    _vec = ()
    _div = ()

    def polys(self):
        """Table slot 0 holds the divergence, slots ``1 + l`` component ``l``
        (the slots an H1 element uses for the value and its derivatives)."""
        return list(self._div), [list(v) for v in self._vec]

    def nbfun_ref(self):
        return len(self._vec)

    def lbasis(self, X, i):
        i = int(i)
        xs = _coords(X, self.dim)
        return ({k: self._vec[i][k](*xs) for k in range(self.dim)},
                self._div[i](*xs))

    def gbasis(self, mapping, X, i, tind):
        if isinstance(X, dict):
            raise NotImplementedError("Calling ElementHdiv gbasis with dict not implemented!")
        x = {k: X[k, :] for k in range(self.dim)}
        phi, dphi = self.lbasis(x, i)
        DF = mapping.DF(X, tind)
        detDF = mapping.detDF(X, tind)
        u = {k: (DF[k][0] * phi[0] + DF[k][1] * phi[1]) / detDF for k in range(2)}
        return u, dphi / detDF


def _p2(terms):
    return Poly(terms, 2)


class ElementTriRT0(ElementHdiv):
    """Lowest order Raviart-Thomas element on triangles, one DOF per facet
    (spfem/element.py:427-452)."""
    maxdeg = 1
    f_dofs = 1
    dim = 2
    _vec = ((_p2({(1, 0): 1}), _p2({(0, 1): 1, (0, 0): -1})),
            (_p2({(1, 0): 1}), _p2({(0, 1): 1})),
            (_p2({(1, 0): -1, (0, 0): 1}), _p2({(0, 1): -1})))
    _div = (_p2({(0, 0): 2}), _p2({(0, 0): 2}), _p2({(0, 0): -2}))


class ElementTriPp(ElementH1):
    """Hierarchical p-basis on triangles (spfem/element.py:623-764): vertex
    functions, ``p-1`` edge functions ``phi_a phi_b iP_j(phi_b - phi_a)`` per
    edge (integrated Legendre polynomials, edges in the mesh's local order
    (0,1), (1,2), (0,2)) and ``(p-1)(p-2)/2`` interior functions
    ``bubble * B_i``.  As in the reference, for ``p > 3`` the ``B_i`` are the
    first ``nbdofs(p-3)`` functions of this very basis.  The polynomials are
    built with exact rational coefficients from the same recurrences."""
    dim = 2

    def __init__(self, p):
        p = int(p)
        self.p = p
        self.maxdeg = p
        self.n_dofs = 1
        self.f_dofs = max(p - 1, 0)
        self.i_dofs = max((p - 1) * (p - 2) // 2, 0)
        self.nbdofs = 3 * self.n_dofs + 3 * self.f_dofs + self.i_dofs
        self._polys = self._build(p)

    @staticmethod
    def _intleg(eta, n):
        """Integrated Legendre polynomials ``iP_0 .. `` in the linear polynomial
        ``eta`` (spfem/element.py:638-663 with its argument ``n`` = p - 2)."""
        n = n + 1
        one = Poly({(0, 0): 1}, 2)
        P = {0: one}
        if n > 1:
            P[1] = eta
        for i in range(1, n):
            P[i + 1] = (eta * P[i]).scale(Fraction(2 * i + 1, i + 1)) - P[i - 1].scale(Fraction(i, i + 1))
        iP = {0: one}
        if n > 1:
            iP[1] = eta
        for i in range(1, n - 1):
            iP[i + 1] = (P[i + 1] - P[i - 1]).scale(Fraction(1, 2 * i + 1))
        return [iP[k] for k in range(len(iP))]

    @classmethod
    def _build(cls, p):
        phi = list(ElementTriP1._polys)
        out = list(phi)
        if p > 1:
            for a, b in ((0, 1), (1, 2), (0, 2)):
                for Pj in cls._intleg(phi[b] - phi[a], p - 2):
                    out.append(phi[a] * phi[b] * Pj)
        if p > 2:
            if p > 3:
                pm3 = p - 3
                N = 3 + 3 * max(pm3 - 1, 0) + max((pm3 - 1) * (pm3 - 2) // 2, 0)
                B = out[:N]
            else:
                B = [Poly({(0, 0): 1}, 2)]
            bubble = phi[0] * phi[1] * phi[2]
            for Bi in B:
                out.append(bubble * Bi)
        return out


class ElementTriDG(ElementH1):
    """Discontinuous version of a triangular H1 element: every DOF becomes an
    interior DOF (spfem/element.py:766-777)."""

    def __init__(self, elem):
        self.elem = elem
        self.maxdeg = elem.maxdeg
        self.i_dofs = 3 * elem.n_dofs + 3 * elem.f_dofs + elem.i_dofs
        self.dim = 2

    def polys(self):
        return self.elem.polys()

    def nbfun_ref(self):
        return self.elem.nbfun_ref()

    def lbasis(self, X, i):
        return self.elem.lbasis(X, i)


class ElementTetDG(ElementH1):
    """Discontinuous version of a tetrahedral H1 element
    (spfem/element.py:779-789)."""

    def __init__(self, elem):
        self.elem = elem
        self.maxdeg = elem.maxdeg
        self.i_dofs = (4 * elem.n_dofs + 4 * elem.f_dofs + 6 * elem.e_dofs
                       + elem.i_dofs)
        self.dim = 3

    def polys(self):
        return self.elem.polys()

    def nbfun_ref(self):
        return self.elem.nbfun_ref()

    def lbasis(self, X, i):
        return self.elem.lbasis(X, i)


def scalar_parent(elem):
    """``(scalar element providing the polynomials, number of components)``."""
    if isinstance(elem, ElementH1Vec):
        return elem.elem, elem.dim
    return elem, 1


def tabulate(elem, X):
    """Reference tables ``phi[i, q]``, ``dphi[i, k, q]`` of the *scalar*
    parent at the points ``X`` (dim x nqp)."""
    base, _ = scalar_parent(elem)
    P, dP = base.polys()
    dim = X.shape[0]
    xs = [X[k, :] for k in range(dim)]
    phi = np.array([p(*xs) for p in P])
    dphi = np.array([[dp[k](*xs) for k in range(dim)] for dp in dP])
    return phi, dphi
