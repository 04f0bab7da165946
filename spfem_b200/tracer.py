# -*- coding: utf-8 -*-
"""Symbolic tracing of user weak forms.

The reference evaluates the user's ``form(u, v, du, dv, x, w, h)`` once per
pair of local basis functions on whole-mesh ``(nt, nqp)`` NumPy arrays
(spfem/assembly.py:971-981).  Here the same Python callable is called a handful
of times with *proxy* arguments and the result is reduced to a normal form
that the CUDA code generator can precontract:

    integrand = sum_{(a, b)}  G^{ab}  *  psi^a_j  *  psi^b_i

``psi^0 = phi`` (reference basis value), ``psi^{1+l} = d phi / d X_l``
(reference derivative), ``a``/``b`` = -1 when the trial/test function does not
occur in the term.  The coefficients ``G^{ab}`` are expression DAGs
(:class:`Coef`) over per-element geometry (``invA``, ``|detA|``, ``h``),
per-point data (``x``, ``w``, normals ``n``) and Python constants captured by
the form; they never contain a basis function.  ``du[k]`` is handed to the
form as ``sum_l invDF[l][k] * psi^{1+l}`` -- the push-forward of
spfem/element.py:480-485 -- so the chain rule is part of ``G``.

Forms that are not multilinear in the basis functions (``abs(u)*v``,
``u**2``) cannot be brought to this form and raise :class:`TraceError`.
"""
import math
import numpy as np


class TraceError(Exception):
    """The weak form cannot be traced into the multilinear normal form."""


# ---------------------------------------------------------------------------
# coefficient DAG (hash-consed)
# ---------------------------------------------------------------------------
_UNARY = {"neg", "abs", "sin", "cos", "tan", "exp", "log", "sqrt", "arcsin",
          "arccos", "arctan", "sinh", "cosh", "tanh", "not", "log10", "log2",
          "floor", "ceil", "sign"}
_BINARY = {"add", "sub", "mul", "div", "pow", "lt", "le", "gt", "ge", "eq",
           "ne", "and", "or", "max", "min", "arctan2"}
_COMMUTATIVE = {"add", "mul", "eq", "ne", "and", "or", "max", "min"}

_PYFUN = {
    "neg": lambda a: -a, "abs": abs, "sin": math.sin, "cos": math.cos,
    "tan": math.tan, "exp": math.exp, "log": math.log, "sqrt": math.sqrt,
    "arcsin": math.asin, "arccos": math.acos, "arctan": math.atan,
    "sinh": math.sinh, "cosh": math.cosh, "tanh": math.tanh,
    "log10": math.log10, "log2": math.log2, "floor": math.floor,
    "ceil": math.ceil, "not": lambda a: float(not a),
    "sign": lambda a: float((a > 0) - (a < 0)),
    "add": lambda a, b: a + b, "sub": lambda a, b: a - b,
    "mul": lambda a, b: a * b, "div": lambda a, b: a / b,
    "pow": lambda a, b: float(np.float64(a) ** np.float64(b)),
    "lt": lambda a, b: float(a < b), "le": lambda a, b: float(a <= b),
    "gt": lambda a, b: float(a > b), "ge": lambda a, b: float(a >= b),
    "eq": lambda a, b: float(a == b), "ne": lambda a, b: float(a != b),
    "and": lambda a, b: float(bool(a) and bool(b)),
    "or": lambda a, b: float(bool(a) or bool(b)),
    "max": max, "min": min, "arctan2": math.atan2,
}


class Coef(object):
    """Node of the coefficient DAG.  ``kind`` is ``'const'``, ``'sym'`` or an
    operator name; nodes are interned so identical sub-expressions are the
    same object (this is the CSE the generated code relies on)."""

    __slots__ = ("kind", "args", "value", "qdep", "uid", "is_bool")
    _pool = {}
    _count = 0

    def __init__(self):
        raise RuntimeError("use Coef.const / Coef.sym / Coef.op")

    @classmethod
    def _make(cls, key, kind, args, value, qdep, is_bool=False):
        node = cls._pool.get(key)
        if node is None:
            node = object.__new__(cls)
            node.kind, node.args, node.value = kind, args, value
            node.qdep = qdep
            node.is_bool = is_bool
            node.uid = cls._count
            cls._count += 1
            cls._pool[key] = node
        return node

    @classmethod
    def const(cls, value):
        value = float(value)
        # keep the sign of zero out of the key: 0.0 and -0.0 behave the same
        if value == 0.0:
            value = 0.0
        return cls._make(("const", value.hex()), "const", (), value, False)

    @classmethod
    def sym(cls, name, qdep=False):
        return cls._make(("sym", name), "sym", (), name, qdep)

    @classmethod
    def op(cls, kind, *args):
        args = tuple(args)
        if all(a.kind == "const" for a in args):
            try:
                return cls.const(_PYFUN[kind](*[a.value for a in args]))
            except (ValueError, ZeroDivisionError, OverflowError):
                pass
        if kind in ("add", "sub", "mul", "div", "neg", "pow"):
            s = cls._simplify(kind, args)
            if s is not None:
                return s
        if kind in _COMMUTATIVE and args[0].uid > args[1].uid:
            args = (args[1], args[0])
        is_bool = kind in ("lt", "le", "gt", "ge", "eq", "ne", "and", "or",
                           "not")
        return cls._make((kind,) + tuple(a.uid for a in args), kind, args,
                         None, any(a.qdep for a in args), is_bool)

    @classmethod
    def _simplify(cls, kind, args):
        def isc(a, v):
            return a.kind == "const" and a.value == v
        if kind == "add":
            a, b = args
            if isc(a, 0.0):
                return b
            if isc(b, 0.0):
                return a
        elif kind == "sub":
            a, b = args
            if isc(b, 0.0):
                return a
            if isc(a, 0.0):
                return cls.op("neg", b)
            if a is b:
                return cls.const(0.0)
        elif kind == "mul":
            a, b = args
            if isc(a, 0.0) or isc(b, 0.0):
                return cls.const(0.0)
            if isc(a, 1.0):
                return b
            if isc(b, 1.0):
                return a
            if isc(a, -1.0):
                return cls.op("neg", b)
            if isc(b, -1.0):
                return cls.op("neg", a)
        elif kind == "div":
            a, b = args
            if isc(a, 0.0):
                return cls.const(0.0)
            if isc(b, 1.0):
                return a
        elif kind == "neg":
            (a,) = args
            if a.kind == "neg":
                return a.args[0]
        elif kind == "pow":
            a, b = args
            if isc(b, 1.0):
                return a
            if isc(b, 0.0):
                return cls.const(1.0)
            if isc(b, 2.0):
                return cls.op("mul", a, a)
        return None

    def is_zero(self):
        return self.kind == "const" and self.value == 0.0

    def __repr__(self):
        if self.kind == "const":
            return repr(self.value)
        if self.kind == "sym":
            return self.value
        return "%s(%s)" % (self.kind, ", ".join(map(repr, self.args)))


class FieldContext(object):
    """Host-evaluated coefficient fields (the H3 fallback of SURVEY.md 7 /
    Appendix B).  When a form applies something the tracer cannot follow to
    ``x``, ``h`` or ``w`` -- ``np.interp``, a SymPy ``lambdify`` with
    ``Piecewise``, a matplotlib interpolator (spfem/mesh.py:885-901,
    spfem/test_module.py:759-765) -- it is traced again with those arguments as
    the reference's own ``(elements, quadrature points)`` NumPy arrays
    (spfem/assembly.py:947-959).  Whatever the opaque code computes from them is
    an ordinary array; the moment such an array meets a basis-function proxy it
    becomes a *field*: a point-dependent symbol ``fld<i>`` whose values the
    kernel streams from HBM."""

    MAX_FIELDS = 8

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.arrays = []
        self._index = {}

    def register(self, arr):
        import zlib
        a = np.asarray(arr)
        if a.dtype == object:
            raise TraceError("an object array met a basis function inside a form")
        try:
            b = np.broadcast_to(a.astype(np.float64, copy=False), self.shape)
        except ValueError:
            raise TraceError("an array of shape %s does not broadcast to (elements, "
                             "quadrature points) = %s" % (a.shape, self.shape))
        b = np.ascontiguousarray(b)
        key = (zlib.adler32(memoryview(b).cast("B")), b.shape)
        for idx in self._index.get(key, ()):          # the same field, computed again
            if np.array_equal(self.arrays[idx], b, equal_nan=True):   # (vector elements: one
                return idx                                            # call per component pair)
        if len(self.arrays) >= self.MAX_FIELDS:
            raise TraceError("more than %d distinct host-evaluated coefficient fields in "
                             "one form" % self.MAX_FIELDS)
        self.arrays.append(b)
        self._index.setdefault(key, []).append(len(self.arrays) - 1)
        return len(self.arrays) - 1


_FIELDS = None        # the FieldContext of the trace in progress (field mode only)


def as_coef(x):
    if isinstance(x, Coef):
        return x
    if _FIELDS is not None and isinstance(x, np.ndarray) and x.ndim >= 1 and x.dtype != object:
        return Coef.sym("fld%d" % _FIELDS.register(x), True)
    if isinstance(x, (bool, np.bool_)):
        return Coef.const(1.0 if x else 0.0)
    if isinstance(x, (int, float, np.integer, np.floating)):
        return Coef.const(float(x))
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return Coef.const(float(x))
    raise TraceError("cannot use a value of type %s inside a traced form "
                     "(only Python / NumPy scalars and the form arguments are "
                     "supported)" % type(x).__name__)


# ---------------------------------------------------------------------------
# multilinear polynomial in the basis functions
# ---------------------------------------------------------------------------
NONE = -1  # "no trial / test factor"


class Form(object):
    """``sum_{(a,b)} coef * psi^a(trial) * psi^b(test)``; the proxy object the
    user's form computes with."""

    __slots__ = ("terms",)
    __array_priority__ = 1000.0

    def __init__(self, terms=None):
        self.terms = {}
        if terms:
            for k, c in terms.items():
                if not c.is_zero():
                    self.terms[k] = c

    # -- constructors -----------------------------------------------------
    @staticmethod
    def coef(c):
        return Form({(NONE, NONE): as_coef(c)})

    @staticmethod
    def wrap(x):
        if isinstance(x, Form):
            return x
        return Form.coef(x)

    def is_pure(self):
        return all(k == (NONE, NONE) for k in self.terms)

    def pure(self, what="this operation"):
        if not self.is_pure():
            raise TraceError("%s would be nonlinear in a basis function; only "
                             "forms that are (multi)linear in u/du and v/dv "
                             "can be compiled" % what)
        return self.terms.get((NONE, NONE), Coef.const(0.0))

    # -- ring operations ----------------------------------------------------
    def _addsub(self, other, op):
        other = Form.wrap(other)
        out = dict(self.terms)
        for k, c in other.terms.items():
            if k in out:
                out[k] = Coef.op(op, out[k], c)
            else:
                out[k] = c if op == "add" else Coef.op("neg", c)
        return Form(out)

    def __add__(self, other):
        if isinstance(other, np.ndarray) and other.dtype == object:
            return NotImplemented
        return self._addsub(other, "add")

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, np.ndarray) and other.dtype == object:
            return NotImplemented
        return self._addsub(other, "sub")

    def __rsub__(self, other):
        return Form.wrap(other)._addsub(self, "sub")

    def __neg__(self):
        return Form({k: Coef.op("neg", c) for k, c in self.terms.items()})

    def __pos__(self):
        return self

    def __mul__(self, other):
        if isinstance(other, np.ndarray) and other.dtype == object:
            return NotImplemented
        other = Form.wrap(other)
        out = {}
        for (a1, b1), c1 in self.terms.items():
            for (a2, b2), c2 in other.terms.items():
                if (a1 != NONE and a2 != NONE) or (b1 != NONE and b2 != NONE):
                    raise TraceError("product of two trial (or two test) "
                                     "function factors: the form is not "
                                     "multilinear")
                k = (a1 if a1 != NONE else a2, b1 if b1 != NONE else b2)
                c = Coef.op("mul", c1, c2)
                out[k] = Coef.op("add", out[k], c) if k in out else c
        return Form(out)

    __rmul__ = __mul__

    def __truediv__(self, other):
        d = Form.wrap(other).pure("division by a basis function")
        return Form({k: Coef.op("div", c, d) for k, c in self.terms.items()})

    def __rtruediv__(self, other):
        d = self.pure("division by a basis function")
        return Form.wrap(other) / Form.coef(d)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __pow__(self, other):
        e = Form.wrap(other).pure("a basis function in an exponent")
        if e.kind == "const" and e.value == 1.0:
            return self
        if self.is_pure():
            return Form.coef(Coef.op("pow", self.pure(), e))
        if e.kind == "const" and e.value == 0.0:
            return Form.coef(1.0)
        raise TraceError("power of a basis function: the form is not "
                         "multilinear")

    def __rpow__(self, other):
        e = self.pure("a basis function in an exponent")
        return Form.coef(Coef.op("pow", as_coef(other), e))

    def __abs__(self):
        return Form.coef(Coef.op("abs", self.pure("abs()")))

    # -- comparisons / boolean algebra (masks such as (x[0] == 1)*v) ----------
    def _cmp(self, other, op):
        return Form.coef(Coef.op(op, self.pure("a comparison"),
                                 Form.wrap(other).pure("a comparison")))

    def __lt__(self, o):
        return self._cmp(o, "lt")

    def __le__(self, o):
        return self._cmp(o, "le")

    def __gt__(self, o):
        return self._cmp(o, "gt")

    def __ge__(self, o):
        return self._cmp(o, "ge")

    def __eq__(self, o):
        return self._cmp(o, "eq")

    def __ne__(self, o):
        return self._cmp(o, "ne")

    __hash__ = None

    def _logic(self, other, op):
        a = self.pure("a boolean operation")
        b = Form.wrap(other).pure("a boolean operation")
        if not a.is_bool:
            a = Coef.op("ne", a, Coef.const(0.0))
        if not b.is_bool:
            b = Coef.op("ne", b, Coef.const(0.0))
        return Form.coef(Coef.op(op, a, b))

    def __and__(self, o):
        return self._logic(o, "and")

    __rand__ = __and__

    def __or__(self, o):
        return self._logic(o, "or")

    __ror__ = __or__

    def __invert__(self):
        a = self.pure("a boolean operation")
        if not a.is_bool:
            a = Coef.op("ne", a, Coef.const(0.0))
        return Form.coef(Coef.op("not", a))

    def __bool__(self):
        raise TraceError("the truth value of a form argument was requested "
                         "(data-dependent Python control flow cannot be "
                         "compiled); use masks such as (x[0] > 0)*v instead")

    # -- NumPy protocol -------------------------------------------------------
    _UFUNC = {
        "add": "add", "subtract": "sub", "multiply": "mul",
        "true_divide": "div", "divide": "div", "power": "pow",
        "negative": "neg", "positive": "pos", "absolute": "abs",
        "fabs": "abs", "sin": "sin", "cos": "cos", "tan": "tan",
        "exp": "exp", "log": "log", "sqrt": "sqrt", "arcsin": "arcsin",
        "arccos": "arccos", "arctan": "arctan", "sinh": "sinh",
        "cosh": "cosh", "tanh": "tanh", "log10": "log10", "log2": "log2",
        "floor": "floor", "ceil": "ceil", "sign": "sign", "square": "square",
        "less": "lt", "less_equal": "le", "greater": "gt",
        "greater_equal": "ge", "equal": "eq", "not_equal": "ne",
        "logical_and": "and", "logical_or": "or", "logical_not": "not",
        "bitwise_and": "and", "bitwise_or": "or", "invert": "not",
        "maximum": "max", "minimum": "min", "arctan2": "arctan2",
        "reciprocal": "recip",
    }

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        if any(isinstance(i, np.ndarray) and i.dtype == object for i in inputs):
            return NotImplemented
        op = Form._UFUNC.get(ufunc.__name__)
        if op is None:
            raise TraceError("NumPy function %s is not supported inside a "
                             "compiled form" % ufunc.__name__)
        x = [Form.wrap(i) for i in inputs]
        if op == "add":
            return x[0] + x[1]
        if op == "sub":
            return x[0] - x[1]
        if op == "mul":
            return x[0] * x[1]
        if op == "div":
            return x[0] / x[1]
        if op == "pow":
            return x[0] ** x[1]
        if op == "neg":
            return -x[0]
        if op == "pos":
            return x[0]
        if op == "square":
            return x[0] * x[0]
        if op == "recip":
            return 1.0 / x[0]
        if op in ("and", "or"):
            return x[0]._logic(x[1], op)
        if op == "not":
            return ~x[0]
        if op in ("lt", "le", "gt", "ge", "eq", "ne"):
            return x[0]._cmp(x[1], op)
        what = "np.%s()" % ufunc.__name__
        return Form.coef(Coef.op(op, *[xi.pure(what) for xi in x]))

    def __float__(self):
        raise TraceError("a form argument was converted to float; call NumPy "
                         "functions (np.sin, np.sqrt, ...) instead of math.*")

    def __repr__(self):
        return "Form(%r)" % (self.terms,)


def zero():
    return Form()


def finalize(result):
    """Normalise what the user's form returned into a :class:`Form`."""
    if isinstance(result, Form):
        return result
    if isinstance(result, np.ndarray):
        if result.dtype == object and result.size == 1:
            return finalize(result.reshape(-1)[0])
        if _FIELDS is not None and result.dtype != object and result.ndim >= 1:
            return Form.coef(result)          # a form without basis functions (norms)
        if result.size == 1:
            return Form.coef(float(result.reshape(-1)[0]))
        raise TraceError("the form returned an array of shape %s; a scalar "
                         "integrand is required" % (result.shape,))
    return Form.coef(result)


# ---------------------------------------------------------------------------
# proxies for the basis functions
# ---------------------------------------------------------------------------
def basis_value(side):
    """``u`` (side 0, trial) or ``v`` (side 1, test) of a scalar element."""
    return Form({((0, NONE) if side == 0 else (NONE, 0)): Coef.const(1.0)})


def basis_grad(side, dim, invdf):
    """Global gradient ``{k: sum_l invDF[l][k] * dphi_l}`` as a real dict
    (spfem/element.py:480-485); ``invdf[l][k]`` are coefficient nodes."""
    out = {}
    for k in range(dim):
        terms = {}
        for l in range(dim):
            key = (1 + l, NONE) if side == 0 else (NONE, 1 + l)
            terms[key] = invdf[l][k]
        out[k] = Form(terms)
    return out


def basis_args(elem_ncomp, comp, side, dim, invdf):
    """``(value, gradient)`` proxies for one local basis function.  Scalar
    elements: a Form and a dict of Forms.  Vector elements
    (``ElementH1Vec``, spfem/element.py:505-546): dicts over the components in
    which only component ``comp`` is non-zero."""
    if elem_ncomp == 1:
        return basis_value(side), basis_grad(side, dim, invdf)
    val, grad = {}, {}
    for c in range(elem_ncomp):
        if c == comp:
            val[c] = basis_value(side)
            grad[c] = basis_grad(side, dim, invdf)
        else:
            val[c] = zero()
            grad[c] = {k: zero() for k in range(dim)}
    return val, grad


def hdiv_args(side, dim, df, rdet):
    """``(u, div u)`` proxies of an H(div) basis function under the Piola map
    (spfem/element.py:407-425): ``u[k] = sum_l DF[k][l] phi_hat_l / detDF`` as a
    real dict, ``du = div phi_hat / detDF``.  Table slot 0 is the reference
    divergence, slot ``1 + l`` component ``l`` (``ElementHdiv.polys``);
    ``df[k][l]`` and ``rdet`` (= 1/detDF) are coefficient nodes."""
    val = {}
    for k in range(dim):
        terms = {}
        for l in range(dim):
            key = (1 + l, NONE) if side == 0 else (NONE, 1 + l)
            terms[key] = Coef.op("mul", df[k][l], rdet)
        val[k] = Form(terms)
    div = Form({((0, NONE) if side == 0 else (NONE, 0)): rdet})
    return val, div


def zero_like(elem_ncomp, dim):
    """The all-zero ``(value, gradient)`` pair used for the "other side" of an
    interior facet (spfem/assembly.py:1108-1111)."""
    if elem_ncomp == 1:
        return zero(), {k: zero() for k in range(dim)}
    return ({c: zero() for c in range(elem_ncomp)},
            {c: {k: zero() for k in range(dim)} for c in range(elem_ncomp)})


def call_form(form, argnames, available):
    """Call ``form`` with the arguments it names (the by-name matching of
    ``Assembler.fillargs``, spfem/assembly.py:168-254).  A name the assembly
    routine does not provide raises ``IndexError`` like the reference."""
    args = []
    for name in argnames:
        if name not in available:
            raise IndexError("list index out of range (the form asks for the "
                             "unsupported argument %r)" % name)
        args.append(available[name])
    return finalize(form(*args))
