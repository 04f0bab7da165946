# -*- coding: utf-8 -*-
"""Element / facet assembly on the GPU behind the sp.fem API.

Drop-in for the hot path of spfem/assembly.py: ``AssemblerElement.__init__``
(:813-837), ``.iasm`` (:839-1006), ``.fasm`` (:1008-1183) and ``Dofnum``
(:1487-1566).  The call signatures, return types (``scipy.sparse.csr_matrix``
with rows = test DOFs, columns = trial DOFs, explicit zeros kept, sorted
indices; or a 1-D ``ndarray`` for linear forms) and the exceptions are the
reference's.  What happens inside is not:

1. the user's form is *traced* with proxy arguments (``tracer.py``),
2. CUDA C for one fused kernel is generated (``codegen.py``) and compiled with
   NVRTC for sm_100a through the C-ABI library (``_lib.py`` ->
   ``csrc/spfem_b200.cu``),
3. the sparsity pattern -- CSR ``indptr``/``indices`` plus the map from CSR
   slots to the local-matrix entries that are summed into them -- is built on
   the device once per assembler (radix sort + run-length encoding),
4. the kernel writes the local matrices, and a deterministic, fixed-order
   gather-reduction sums them into the CSR values (no floating point atomics).

There is no CPU fallback: without the CUDA library / a GPU the calls raise.
"""
import inspect
import os
import numpy as np

from . import mesh as _mesh
from . import element as _element
from .quadrature import get_quadrature
from .tracer import (Coef, Form, TraceError, basis_args, hdiv_args, zero_like,  # noqa: F401
                     call_form)
from . import codegen
from . import tracer as _tracer

__all__ = ["Assembler", "AssemblerElement", "Dofnum", "TraceError"]


class Dofnum(object):
    """Global degree-of-freedom numbering (spfem/assembly.py:1487-1566):
    vertex DOFs ``v*n_dofs + k`` first, then edge DOFs (3-D meshes), facet
    DOFs, interior DOFs; ``t_dof`` lists, per element, the vertex DOFs of
    each local vertex, then local edges, local facets and interior DOFs."""

    def __init__(self, mesh, element):
        nv = mesh.p.shape[1]
        nt = mesh.t.shape[1]
        self.n_dof = np.arange(element.n_dofs * nv, dtype=np.int64).reshape(
            (element.n_dofs, nv), order='F')
        offset = element.n_dofs * nv
        self.e_dof = np.array([])
        self.f_dof = np.array([])
        has_edges = hasattr(mesh, 'edges')
        has_facets = hasattr(mesh, 'facets')
        if has_edges:
            ne = mesh.edges.shape[1]
            self.e_dof = np.arange(element.e_dofs * ne, dtype=np.int64).reshape(
                (element.e_dofs, ne), order='F') + offset
            offset += element.e_dofs * ne
        if has_facets:
            nf = mesh.facets.shape[1]
            self.f_dof = np.arange(element.f_dofs * nf, dtype=np.int64).reshape(
                (element.f_dofs, nf), order='F') + offset
            offset += element.f_dofs * nf
        self.i_dof = np.arange(element.i_dofs * nt, dtype=np.int64).reshape(
            (element.i_dofs, nt), order='F') + offset
        rows = [self.n_dof[:, mesh.t[k, :]] for k in range(mesh.t.shape[0])]
        if has_edges:
            rows += [self.e_dof[:, mesh.t2e[k, :]] for k in range(mesh.t2e.shape[0])]
        if has_facets:
            rows += [self.f_dof[:, mesh.t2f[k, :]] for k in range(mesh.t2f.shape[0])]
        rows.append(self.i_dof)
        self.t_dof = np.vstack(rows).astype(np.int64)
        self.N = int(np.max(self.t_dof)) + 1

    def getdofs(self, N=None, F=None, E=None, T=None):
        """Global DOF numbers of the given nodes / facets / edges / elements."""
        dofs = np.zeros(0, dtype=np.int64)
        if N is not None:
            dofs = np.hstack((dofs, self.n_dof[:, N].flatten()))
        if F is not None:
            dofs = np.hstack((dofs, self.f_dof[:, F].flatten()))
        if E is not None:
            dofs = np.hstack((dofs, self.e_dof[:, E].flatten()))
        if T is not None:
            dofs = np.hstack((dofs, self.i_dof[:, T].flatten()))
        return dofs.flatten()


class Assembler(object):
    """Base class (spfem/assembly.py:31-254)."""

    def __init__(self):
        raise NotImplementedError("Assembler constructor not implemented!")


def _argnames(form):
    return list(inspect.getfullargspec(form).args)


class _Plan(object):
    """A traced + generated kernel, independent of any device."""

    def __init__(self, kind, spec, source, name, bilinear, wkeys):
        self.kind, self.spec, self.source, self.name = kind, spec, source, name
        self.bilinear, self.wkeys = bilinear, wkeys


class AssemblerElement(Assembler):
    """Finite element assembler for elements defined through reference basis
    functions (spfem/assembly.py:789-837)."""

    def __init__(self, mesh, elem_u, elem_v=None, mapping=None):
        if not isinstance(mesh, _mesh.Mesh):
            raise Exception("First parameter must be an instance of "
                            "spfem.mesh.Mesh!")
        if not isinstance(elem_u, _element.Element):
            raise Exception("Second parameter must be an instance of "
                            "spfem.element.Element!")
        default = mesh.mapping()
        def same_mesh(mp):
            if getattr(mp, "_mesh", None) is mesh:                   # MappingAffine
                return True
            return getattr(mp, "p", None) is mesh.p and getattr(mp, "t", None) is mesh.t
        if mapping is not None and not (type(mapping) is type(default) and same_mesh(mapping)):
            # the device kernels rebuild the geometry from mesh.p / mesh.t with the
            # mesh's default map (spfem/assembly.py:829-832 stores any mapping and
            # uses it for F / invDF / detDF): a different map would be ignored silently
            raise NotImplementedError(
                "AssemblerElement: only the mesh's own mapping (mesh.mapping()) is "
                "supported on the device path; deform mesh.p instead of passing a "
                "pre-deformed mapping")
        self.mapping = default if mapping is None else mapping
        self.mesh = mesh
        self.elem_u = elem_u
        self.dofnum_u = Dofnum(mesh, elem_u)
        if elem_v is None:
            self.elem_v = elem_u
            self.dofnum_v = self.dofnum_u
        else:
            self.elem_v = elem_v
            self.dofnum_v = Dofnum(mesh, elem_v)
        self._engine = None        # device state, created on first assembly
        self.row_mask = None       # optional bool [dofnum_v.N]: rows assembled here (multi-GPU)
        self._pou = None
        self.last_stats = {}

    # ------------------------------------------------------------------
    # tracing
    # ------------------------------------------------------------------
    def _elem_info(self):
        bu, ncu = _element.scalar_parent(self.elem_u)
        bv, ncv = _element.scalar_parent(self.elem_v)
        return bu, ncu, bv, ncv

    def _check_interp(self, interp):
        if interp is None:
            return []
        if not isinstance(interp, dict):
            raise Exception("The input solution vector(s) must be in a "
                            "dictionary! Pass e.g. {0:u} instead of u.")
        keys = list(interp.keys())
        if len(keys) > 8:
            raise NotImplementedError("at most 8 interpolated fields")
        return keys

    def plan_iasm(self, form, intorder=None, interp=None):
        """Trace ``form`` and generate the interior kernel (no device work).

        A form the tracer cannot follow symbolically (it applies an opaque
        callable -- ``np.interp``, a SymPy ``lambdify`` with ``Piecewise``, a
        matplotlib interpolator, spfem/mesh.py:885-901 -- to ``x``, ``h`` or
        ``w``) is traced a second time in *field mode*: those arguments are the
        reference's own ``(elements, quadrature points)`` NumPy arrays
        (spfem/assembly.py:947-959), whatever the form computes from them is
        evaluated on the host exactly as the reference does and streamed to the
        kernel as coefficient fields; ``u, v, du, dv`` stay symbolic.  Forms that
        are not multilinear in the basis functions still raise ``TraceError``."""
        try:
            return self._plan_iasm(form, intorder, interp, False)
        except TraceError as exc:
            if os.environ.get("SPFEM_FIELDS", "1") == "0":
                raise
            try:
                return self._plan_iasm(form, intorder, interp, True)
            except Exception as exc2:
                # neither symbolically nor on the reference's arrays: report the
                # tracer's diagnosis, chained to what the array evaluation said
                raise exc from exc2

    def _host_point_data(self, X, interp, wkeys, bu):
        """``x``, ``h``, ``w`` as the reference evaluates them on the host
        (spfem/assembly.py:947-959, 974): arrays of shape (nt, nq)."""
        mp = self.mapping
        x = mp.F(X)
        det = mp.detDF(X)
        dim = self.mesh.p.shape[0]
        h = np.abs(det) ** (1.0 / dim)
        w = {}
        if wkeys:
            phi, _ = _element.tabulate(bu, X)           # (nbfun, nq) of the scalar parent
            td = self.dofnum_u.t_dof
            for key in wkeys:
                vec = np.asarray(interp[key], dtype=np.float64)
                acc = 0.0 * x[0]
                for j in range(td.shape[0]):
                    acc = acc + np.outer(vec[td[j, :]], phi[j])
                w[key] = acc
        return x, h, w

    def _plan_iasm(self, form, intorder, interp, field_mode):
        mesh = self.mesh
        if intorder is None:
            intorder = self.elem_u.maxdeg + self.elem_v.maxdeg
        names = _argnames(form)
        bilinear = ('u' in names) or ('du' in names)
        wkeys = self._check_interp(interp)
        X, W = get_quadrature(mesh.refdom, intorder)
        dim = mesh.p.shape[0]
        bu, ncu, bv, ncv = self._elem_info()
        if not (hasattr(bu, "polys") and hasattr(bv, "polys")):
            raise NotImplementedError("element type without reference "
                                      "polynomials is not supported on the "
                                      "device path")
        quad = mesh.refdom == "quad"
        if mesh.refdom not in ("tri", "tet", "quad"):
            raise NotImplementedError("iasm: unsupported reference domain %r"
                                      % (mesh.refdom,))
        if quad:
            invdf = {l: {k: Coef.sym("iJ%d%d" % (l, k), True) for k in range(dim)}
                     for l in range(dim)}
            absdet = Coef.sym("absdetq", True)
            h = Form.coef(Coef.sym("hq", True))
        else:
            invdf = {l: {k: Coef.sym("iA%d%d" % (l, k)) for k in range(dim)}
                     for l in range(dim)}
            absdet = Coef.sym("absdet")
            h = Form.coef(Coef.sym("h"))
        x = {k: Form.coef(Coef.sym("x%d" % k, True)) for k in range(dim)}
        w = {key: Form.coef(Coef.sym("w%d" % s, True))
             for s, key in enumerate(wkeys)}
        fctx = None
        if field_mode:
            if ncu > 1 and wkeys:
                raise TraceError("field mode: interpolated fields of vector elements")
            x, h, w = self._host_point_data(X, interp, wkeys, bu)
            fctx = _tracer.FieldContext((mesh.t.shape[1], X.shape[1]))
        hdiv_u, hdiv_v = getattr(bu, "hdiv", False), getattr(bv, "hdiv", False)
        if hdiv_u or hdiv_v:
            if quad or dim != 2:
                raise NotImplementedError("H(div) elements: affine triangles only")
            df = {k: {l: Coef.sym("A_%d%d" % (k, l)) for l in range(dim)} for k in range(dim)}
            rdet = Coef.op("div", Coef.const(1.0), Coef.sym("detA"))
        blocks = {}
        _tracer._FIELDS = fctx
        try:
            for cu in range(ncu if bilinear else 1):
                for cv in range(ncv):
                    v, dv = (hdiv_args(1, dim, df, rdet) if hdiv_v
                             else basis_args(ncv, cv, 1, dim, invdf))
                    avail = {'v': v, 'dv': dv, 'x': x, 'w': w, 'h': h}
                    if bilinear:
                        u, du = (hdiv_args(0, dim, df, rdet) if hdiv_u
                                 else basis_args(ncu, cu, 0, dim, invdf))
                        avail['u'] = u
                        avail['du'] = du
                    try:
                        f = call_form(form, names, avail)
                    except TraceError:
                        raise
                    except (TypeError, ValueError, AttributeError, NotImplementedError) as exc:
                        if field_mode:
                            raise
                        # an opaque callable choked on the symbolic proxies
                        raise TraceError("the form could not be traced symbolically (%s: %s)"
                                         % (type(exc).__name__, exc))
                    # the Jacobian factor |detDF| of spfem/assembly.py:980-981
                    blocks[(cu, cv)] = f * Form.coef(absdet)
        finally:
            _tracer._FIELDS = None
        tab_u = self._tables(bu, X)
        tab_v = self._tables(bv, X)
        nfld = len(fctx.arrays) if fctx is not None else 0
        spec = codegen.InteriorSpec(mesh.refdom, dim, bilinear,
                                    bu.nbfun_ref(), bv.nbfun_ref(), ncu, ncv,
                                    blocks, X, W, tab_u, tab_v,
                                    nw=0 if field_mode else len(wkeys), nfld=nfld)
        src = codegen.generate_interior(spec, "spfem_elem")
        plan = _Plan("interior", spec, src, "spfem_elem", bilinear,
                     [] if field_mode else wkeys)
        plan.fields = fctx.arrays if fctx is not None else []
        return plan

    @staticmethod
    def _tables(base, X):
        phi, dphi = _element.tabulate(base, X)
        dim = X.shape[0]
        return [phi] + [dphi[:, l, :] for l in range(dim)]

    def plan_fasm(self, form, interior=False, intorder=None, normals=True,
                  interp=None, find=None):
        """Trace ``form`` and generate the facet kernel (no device work).  Like
        :meth:`plan_iasm`, a form that applies opaque callables to ``x``, ``h`` or
        ``n`` is traced a second time on the reference's own ``(facets, points)``
        arrays (spfem/assembly.py:1046-1075) for the facets ``find``."""
        try:
            return self._plan_fasm(form, interior, intorder, normals, interp, None)
        except TraceError as exc:
            if os.environ.get("SPFEM_FIELDS", "1") == "0":
                raise
            if find is None:
                find = (self.mesh.interior_facets() if interior
                        else self.mesh.boundary_facets())
            try:
                return self._plan_fasm(form, interior, intorder, normals, interp,
                                       np.asarray(find))
            except Exception as exc2:
                raise exc from exc2

    def _plan_fasm(self, form, interior, intorder, normals, interp, field_find):
        mesh = self.mesh
        if intorder is None:
            intorder = self.elem_u.maxdeg + self.elem_v.maxdeg
        names = _argnames(form)
        if interior:
            bilinear = ('u1' in names) or ('du1' in names)
        else:
            bilinear = ('u' in names) or ('du' in names)
        wkeys = self._check_interp(interp)
        if not bilinear and interior:
            raise Exception("No interior support in linear facet form.")
        X, W = get_quadrature(mesh.brefdom, intorder)
        X = np.atleast_2d(X)
        dim = mesh.p.shape[0]
        if mesh.refdom not in ("tri", "tet", "quad"):
            raise NotImplementedError("fasm: unsupported reference domain %r"
                                      % (mesh.refdom,))
        if mesh.refdom == "quad" and (interior or normals):
            raise NotImplementedError("MappingQ1 has no normals / interior "
                                      "facet support (pass normals=False)")
        bu, ncu, bv, ncv = self._elem_info()
        inv = {s: {l: {k: Coef.sym("iA%d_%d%d" % (s, l, k), True)
                       for k in range(dim)} for l in range(dim)}
               for s in (1, 2)}
        x = {k: Form.coef(Coef.sym("x%d" % k, True)) for k in range(dim)}
        h = Form.coef(Coef.sym("hf", True))
        n = ({k: Form.coef(Coef.sym("n%d" % k, True)) for k in range(dim)}
             if normals else {})
        w = {key: Form.coef(Coef.sym("w%d" % s, True))
             for s, key in enumerate(wkeys)}
        dw = {key: {a: Form.coef(Coef.sym("dw%d_%d" % (s, a), True))
                    for a in range(dim)} for s, key in enumerate(wkeys)}
        fctx = None
        if field_find is not None:
            if wkeys:
                raise TraceError("field mode: interpolated fields in facet forms")
            if mesh.refdom == "quad":
                raise TraceError("field mode: facet forms on quadrilaterals")
            # x, h, n as the reference evaluates them   (spfem/assembly.py:1046-1075)
            mp = self.mapping
            nq = X.shape[1]
            x = mp.G(X, field_find)
            detB = np.tile(mp.detB[field_find][:, None], (1, nq))
            h = np.abs(detB) ** (1.0 / (dim - 1.0))
            n = (mp.normals(X, mesh.f2t[0, field_find], field_find, mesh.t2f)
                 if normals else {})
            fctx = _tracer.FieldContext((len(field_find), nq))
        sides = [(1, 1), (2, 2), (2, 1), (1, 2)] if interior else [(1, 1)]
        blocks = []
        _tracer._FIELDS = fctx
        for (su, sv) in sides:
            blk = {}
            for cu in range(ncu if bilinear else 1):
                for cv in range(ncv):
                    avail = {'x': x, 'h': h, 'n': n, 'w': w, 'dw': dw}
                    if interior:
                        zu, dzu = zero_like(ncu, dim)
                        zv, dzv = zero_like(ncv, dim)
                        ub, dub = basis_args(ncu, cu, 0, dim, inv[su])
                        vb, dvb = basis_args(ncv, cv, 1, dim, inv[sv])
                        avail.update({
                            'u1': ub if su == 1 else zu, 'u2': ub if su == 2 else zu,
                            'du1': dub if su == 1 else dzu, 'du2': dub if su == 2 else dzu,
                            'v1': vb if sv == 1 else zv, 'v2': vb if sv == 2 else zv,
                            'dv1': dvb if sv == 1 else dzv, 'dv2': dvb if sv == 2 else dzv})
                    else:
                        v, dv = basis_args(ncv, cv, 1, dim, inv[1])
                        avail.update({'v': v, 'dv': dv})
                        if bilinear:
                            u, du = basis_args(ncu, cu, 0, dim, inv[1])
                            avail.update({'u': u, 'du': du})
                    try:
                        blk[(cu, cv)] = call_form(form, names, avail)
                    except TraceError:
                        _tracer._FIELDS = None
                        raise
                    except (TypeError, ValueError, AttributeError, NotImplementedError) as exc:
                        _tracer._FIELDS = None
                        if field_find is not None:
                            raise
                        raise TraceError("the form could not be traced symbolically (%s: %s)"
                                         % (type(exc).__name__, exc))
            blocks.append(blk)
        _tracer._FIELDS = None
        spec = codegen.FacetSpec(mesh.refdom, dim, bilinear, interior,
                                 bu.nbfun_ref(), bv.nbfun_ref(), ncu, ncv,
                                 blocks, sides, X, W, bu.polys(), bv.polys(),
                                 normals=normals, nw=len(wkeys),
                                 nfld=len(fctx.arrays) if fctx is not None else 0)
        src = codegen.generate_facet(spec, "spfem_facet")
        plan = _Plan("facet", spec, src, "spfem_facet", bilinear, wkeys)
        plan.fields = fctx.arrays if fctx is not None else []
        return plan

    # ------------------------------------------------------------------
    # index tables shared by the device path and the tests
    # ------------------------------------------------------------------
    def facet_tables(self, find):
        """``(fv, e1, e2, lf1)`` for the facets ``find``: facet vertex ids,
        the two neighbouring elements (``f2t``, second = -1 on the boundary)
        and the local index of the facet inside the first element (used for
        the reference normal, spfem/mapping.py:494-497)."""
        mesh = self.mesh
        find = np.asarray(find, dtype=np.int64)
        fv = mesh.facets[:, find]
        e1 = mesh.f2t[0, find]
        e2 = mesh.f2t[1, find]
        lf1 = np.zeros(find.shape[0], dtype=np.int64)
        for itr in range(mesh.t2f.shape[0]):
            lf1[mesh.t2f[itr, e1] == find] = itr
        return fv, e1, e2, lf1

    def facet_dofs(self, find, interior):
        """Row / column DOF tables of the facet COO blocks
        (spfem/assembly.py:1131-1156)."""
        e1 = self.mesh.f2t[0, find]
        if not interior:
            return self.dofnum_v.t_dof[:, e1], self.dofnum_u.t_dof[:, e1]
        e2 = self.mesh.f2t[1, find]
        tv, tu = self.dofnum_v.t_dof, self.dofnum_u.t_dof
        rows = np.hstack((tv[:, e1], tv[:, e2], tv[:, e1], tv[:, e2]))
        cols = np.hstack((tu[:, e1], tu[:, e2], tu[:, e2], tu[:, e1]))
        return rows, cols

    # ------------------------------------------------------------------
    # public API
    # ------------------------------------------------------------------
    def _get_engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        return self._engine

    def iasm(self, form, intorder=None, tind=None, interp=None):
        """Assemble a bilinear or linear form over element interiors.

        Same contract as spfem/assembly.py:839-1006: the form's argument names
        select among ``u, v, du, dv, x, w, h``; bilinear iff it names ``u`` or
        ``du``.  Returns a CSR matrix of shape ``(dofnum_v.N, dofnum_u.N)`` or
        a vector of length ``dofnum_v.N``."""
        plan = self.plan_iasm(form, intorder=intorder, interp=interp)
        chunks = self._chunks_needed(plan, tind)
        if chunks > 1:
            if tind is None and "SPFEM_MAX_ENTRIES" not in os.environ:
                # the native tile plan is built batch by batch and has no 2^31 limit:
                # one fused launch per batch; only forms that need the two-kernel
                # route fall back to the element-chunk loop
                try:
                    return self._get_engine().run_interior(plan, None, interp)
                except NotImplementedError:
                    pass
            return self._iasm_chunked(form, plan, intorder, interp, chunks)
        return self._get_engine().run_interior(plan, tind, interp)

    # -- meshes beyond one pattern call -----------------------------------
    def _chunks_needed(self, plan, tind):
        """Number of element chunks so that one chunk (with its halo) stays
        below the limit of local entries per pattern call (2^31 - 1;
        ``SPFEM_MAX_ENTRIES`` lowers it for the tests)."""
        limit = int(os.environ.get("SPFEM_MAX_ENTRIES", str(2 ** 31 - 1)))
        nt = self.mesh.t.shape[1] if tind is None else len(np.atleast_1d(tind))
        n_ent = plan.spec.n_entries * nt
        if n_ent <= limit:
            return 1
        if tind is not None:
            raise NotImplementedError("tind subsets with more than %d local entries" % limit)
        return int(np.ceil(1.4 * n_ent / limit))

    def _iasm_chunked(self, form, plan, intorder, interp, chunks, assemble=None,
                      to_host=True):
        """Assemble chunk by chunk on one GPU: Morton blocks of elements, every
        global row owned by the lowest chunk touching it, a chunk's local mesh =
        its block plus the halo elements of its rows (the multi-GPU partition of
        ``distributed.py`` run sequentially).  Owned rows are complete after their
        chunk, so the global CSR is the union of the owned rows; pattern and index
        width are those of the single call (SciPy's rule on the *total* number of
        local entries).

        The merge runs on the device: a chunk's owned entries are renumbered to
        global ids and sorted by (row, column) key there and stay in HBM; the rows
        of different chunks are disjoint, so once every chunk has reported its
        row lengths the global ``indptr`` is one prefix sum and each chunk's
        entries are scattered to ``indptr[row] + position in the row`` -- no
        global sort, no host-side COO.  One device-to-host copy at the end."""
        import torch
        from .distributed import Partitioner, _local_tensors
        from .engine import scipy_index_dtype
        from scipy.sparse import csr_matrix
        ev = None if self.elem_v is self.elem_u else self.elem_v
        limit = int(os.environ.get("SPFEM_MAX_ENTRIES", str(2 ** 31 - 1)))
        Nv, Nu = self.dofnum_v.N, self.dofnum_u.N
        bilinear = plan.bilinear
        pieces = []                 # per chunk: (global rows, their lengths, columns, values)
        counts = out = None
        parts = Partitioner(self.mesh, self.elem_u, ev, chunks)
        r = 0
        while r < chunks:
            part = parts.part(r)
            if plan.spec.n_entries * part.mesh.t.shape[1] > limit:
                chunks, r, pieces, counts, out = 2 * chunks, 0, [], None, None
                parts = Partitioner(self.mesh, self.elem_u, ev, chunks)   # halo thicker than assumed
                continue
            loc = AssemblerElement(part.mesh, self.elem_u, ev)
            li = None if interp is None else {k: np.asarray(v)[part.l2g_u]
                                              for k, v in interp.items()}
            if assemble is not None:
                A = assemble(loc, form, intorder=intorder, interp=li)
            else:
                A = loc.iasm_device(form, intorder=intorder, interp=li)
            rows, cols, vals = _local_tensors(A, bilinear)
            dev = vals.device
            keep = torch.from_numpy(part.owned_v).to(dev)[rows]
            grow = torch.from_numpy(np.ascontiguousarray(part.l2g_v, dtype=np.int64)).to(dev)[rows[keep]]
            vals = vals[keep]
            if not bilinear:
                if out is None:
                    out = torch.zeros(Nv, dtype=torch.float64, device=dev)
                out[grow] = vals
            else:
                gcol = torch.from_numpy(np.ascontiguousarray(part.l2g_u, dtype=np.int64)).to(dev)[cols[keep]]
                key, order = torch.sort(grow * Nu + gcol)
                grow = torch.div(key, Nu, rounding_mode="floor")
                gcol = key - grow * Nu
                urows, cnt = torch.unique_consecutive(grow, return_counts=True)
                if counts is None:
                    counts = torch.zeros(Nv, dtype=torch.int64, device=dev)
                counts[urows] = cnt                     # (rows of different chunks are disjoint)
                pieces.append((urows, cnt, gcol, vals[order]))
                del key, order, grow
            del A, rows, cols, keep
            loc._engine = None                          # release the chunk's device memory
            r += 1
        if not bilinear:
            return out.cpu().numpy() if to_host else out
        idt = scipy_index_dtype(plan.spec.n_entries * self.mesh.t.shape[1], Nv, Nu)
        tdt = torch.int64 if idt == np.int64 else torch.int32
        dev = counts.device
        indptr = torch.zeros(Nv + 1, dtype=torch.int64, device=dev)
        indptr[1:] = torch.cumsum(counts, 0)
        nnz = int(indptr[-1])
        indices = torch.empty(nnz, dtype=tdt, device=dev)
        data = torch.empty(nnz, dtype=torch.float64, device=dev)
        while pieces:
            urows, cnt, gcol, vals = pieces.pop()
            first = torch.cumsum(cnt, 0) - cnt           # start of each row inside the chunk
            pos = (torch.arange(gcol.numel(), device=dev, dtype=torch.int64)
                   + torch.repeat_interleave(indptr[urows] - first, cnt))
            indices[pos] = gcol.to(tdt)
            data[pos] = vals
            del urows, cnt, gcol, vals, first, pos
        indptr = indptr.to(tdt)
        if not to_host:
            return indptr, indices, data
        A = csr_matrix((Nv, Nu), dtype=np.float64)
        A.indptr, A.indices, A.data = (indptr.cpu().numpy(), indices.cpu().numpy(),
                                       data.cpu().numpy())
        A.has_sorted_indices = True
        A.has_canonical_format = True
        return A

    def iasm_device(self, form, intorder=None, tind=None, interp=None):
        """Like :meth:`iasm` but leaves the result on the GPU
        (:class:`engine.DeviceCSR` or a device tensor for linear forms)."""
        plan = self.plan_iasm(form, intorder=intorder, interp=interp)
        return self._get_engine().run_interior(plan, tind, interp, to_host=False)

    def prepare_iasm(self, form, intorder=None, tind=None, interp=None):
        """Trace, compile and stage ``form`` once; the returned handle's
        ``launch()`` re-runs the assembly on the device (time stepping,
        benchmarks) and ``result()`` fetches it like :meth:`iasm`."""
        plan = self.plan_iasm(form, intorder=intorder, interp=interp)
        return self._get_engine().prepare_interior(plan, tind, interp)

    def prepare_fasm(self, form, find=None, interior=False, intorder=None,
                     normals=True, interp=None):
        """Staged facet assembly (see :meth:`prepare_iasm`)."""
        if find is None:
            find = (self.mesh.interior_facets() if interior
                    else self.mesh.boundary_facets())
        plan = self.plan_fasm(form, interior=interior, intorder=intorder,
                              normals=normals, interp=interp, find=find)
        return self._get_engine().prepare_facet(plan, np.asarray(find), interior, interp)

    # ------------------------------------------------------------------
    # error norms (spfem/assembly.py:1353-1485) on top of the device path
    # ------------------------------------------------------------------
    def _unit_assembler(self):
        """Assembler with a partition-of-unity element on the same mesh: the sum
        of the load vector of ``g * v`` is the integral of ``g``."""
        if getattr(self, "_pou", None) is None:
            elem = {"tri": _element.ElementTriP1, "tet": _element.ElementTetP1,
                    "quad": _element.ElementQ1}.get(self.mesh.refdom)
            if elem is None:
                raise NotImplementedError("error norms on this mesh type")
            self._pou = AssemblerElement(self.mesh, elem())
        return self._pou

    def inorm(self, form, interp, intorder=None):
        """Element-wise squared L2 norms ``int_K form(u, du, x, h)^2`` of
        interpolated solution vectors (spfem/assembly.py:1273-1351; a posteriori
        estimators).  ``u`` / ``du`` are the dicts of interpolated fields and
        their gradients.  Returns an array with one value per element."""
        if not isinstance(interp, dict):
            raise Exception("The input solution vector(s) must be in a "
                            "dictionary! Pass e.g. {0:u} instead of u.")
        mesh = self.mesh
        if intorder is None:
            intorder = 2 * self.elem_u.maxdeg
        names = _argnames(form)
        wkeys = self._check_interp(interp)
        X, W = get_quadrature(mesh.refdom, intorder)
        dim = mesh.p.shape[0]
        bu, ncu, _, _ = self._elem_info()
        if ncu != 1:
            raise NotImplementedError("inorm: scalar elements only")
        quad = mesh.refdom == "quad"
        absdet = Coef.sym("absdetq", True) if quad else Coef.sym("absdet")
        h = Form.coef(Coef.sym("hq", True)) if quad else Form.coef(Coef.sym("h"))
        x = {k: Form.coef(Coef.sym("x%d" % k, True)) for k in range(dim)}
        w = {key: Form.coef(Coef.sym("w%d" % s, True)) for s, key in enumerate(wkeys)}
        dw = {key: {a: Form.coef(Coef.sym("dw%d_%d" % (s, a), True)) for a in range(dim)}
              for s, key in enumerate(wkeys)}
        val = call_form(form, names, {'u': w, 'du': dw, 'x': x, 'h': h})
        val = Form.coef(val.pure("inorm form"))
        # one constant "test function" per element: the entry is the integral
        from .tracer import NONE
        integrand = val * val * Form.coef(absdet)
        blocks = {(0, 0): Form({(NONE, 0): integrand.pure()})}
        ones = [np.ones((1, len(W)))]
        spec = codegen.InteriorSpec(mesh.refdom, dim, False, bu.nbfun_ref(), 1, 1, 1,
                                    blocks, X, W, self._tables(bu, X), ones, nw=len(wkeys))
        src = codegen.generate_interior(spec, "spfem_elem")
        plan = _Plan("functional", spec, src, "spfem_elem", False, wkeys)
        return self._get_engine().run_functional(plan, interp)

    def fnorm(self, form, interp, intorder=None, interior=False, normals=True):
        """Facet-wise squared L2 norms ``int_F form(...)^2`` of interpolated
        solution vectors (spfem/assembly.py:1185-1271): boundary facets with
        ``form(u, du, x, n, t, h)``, interior facets with
        ``form(u1, u2, du1, du2, x, n, t, h)``.  Returns ``(values, find)``."""
        mesh = self.mesh
        find = mesh.interior_facets() if interior else mesh.boundary_facets()
        if not isinstance(interp, dict):
            raise Exception("The input solution vector(s) must be in a "
                            "dictionary! Pass e.g. {0:u} instead of u.")
        if intorder is None:
            intorder = 2 * self.elem_u.maxdeg
        names = _argnames(form)
        if ('u1' in names or 'du1' in names) and not interior:
            raise Exception("The supplied form contains u1 although "
                            "no interior=True is given.")
        wkeys = self._check_interp(interp)
        X, W = get_quadrature(mesh.brefdom, intorder)
        X = np.atleast_2d(X)
        dim = mesh.p.shape[0]
        if mesh.refdom not in ("tri", "tet"):
            raise NotImplementedError("fnorm: unsupported reference domain %r"
                                      % (mesh.refdom,))
        bu, ncu, _, _ = self._elem_info()
        if ncu != 1:
            raise NotImplementedError("fnorm: scalar elements only")
        x = {k: Form.coef(Coef.sym("x%d" % k, True)) for k in range(dim)}
        h = Form.coef(Coef.sym("hf", True))
        n = ({k: Form.coef(Coef.sym("n%d" % k, True)) for k in range(dim)}
             if normals else {})
        t = ({k: Form.coef(Coef.sym("t%d" % k, True)) for k in range(2)}
             if normals and dim == 2 else {})

        def fields(tag):
            w = {key: Form.coef(Coef.sym("w%d%s" % (s, tag), True))
                 for s, key in enumerate(wkeys)}
            dw = {key: {a: Form.coef(Coef.sym("dw%d%s_%d" % (s, tag, a), True))
                        for a in range(dim)} for s, key in enumerate(wkeys)}
            return w, dw
        w1, dw1 = fields("")
        avail = {'x': x, 'n': n, 't': t, 'h': h}
        if interior:
            w2, dw2 = fields("b")
            avail.update({'u1': w1, 'u2': w2, 'du1': dw1, 'du2': dw2})
        else:
            avail.update({'u': w1, 'du': dw1})
        val = Form.coef(call_form(form, names, avail).pure("fnorm form"))
        from .tracer import NONE
        blocks = [{(0, 0): Form({(NONE, NONE): (val * val).pure()})}]
        spec = codegen.FacetSpec(mesh.refdom, dim, False, interior, bu.nbfun_ref(), 1, 1, 1,
                                 blocks, [(1, 1)], X, W, bu.polys(), bu.polys(),
                                 normals=normals, nw=len(wkeys), functional=True)
        src = codegen.generate_facet(spec, "spfem_facet")
        plan = _Plan("facet_functional", spec, src, "spfem_facet", False, wkeys)
        return self._get_engine().run_facet_functional(plan, np.asarray(find), interp), find

    def L2error(self, uh, exact, intorder=None):
        """Global L2 error ``||exact - uh||_0`` through the identity
        ``(u,u) + (uh,uh) - 2 (u,uh)`` (spfem/assembly.py:1353-1402).  ``exact``
        is called with the coordinate dict ``x`` of form arguments."""
        if self.elem_u.maxdeg != self.elem_v.maxdeg:
            raise NotImplementedError("elem_u.maxdeg must be elem_v.maxdeg "
                                      "when computing errors!")
        if intorder is None:
            intorder = 2 * self.elem_u.maxdeg

        def uv(u, v):
            return u * v

        def fv(v, x):
            return exact(x) * v

        def ff(v, x):
            return exact(x) ** 2 * v
        M = self.iasm(uv)
        f = self.iasm(fv)
        uu = float(np.sum(self._unit_assembler().iasm(ff, intorder=intorder)))
        return np.sqrt(uu + np.dot(uh, M.dot(uh)) - 2. * np.dot(uh, f))

    def H1error(self, uh, dexact, intorder=None):
        """Global H1 seminorm error (spfem/assembly.py:1404-1485); ``dexact`` is
        a dict of callables (one per direction), each called with ``x``."""
        if self.elem_u.maxdeg != self.elem_v.maxdeg:
            raise NotImplementedError("elem_u.maxdeg must be elem_v.maxdeg "
                                      "when computing errors!")
        if intorder is None:
            intorder = 2 * self.elem_u.maxdeg
        dim = self.mesh.p.shape[0]
        if dim not in (2, 3):
            raise NotImplementedError("AssemblerElement.H1error not "
                                      "implemented for current domain "
                                      "dimension!")

        def uv(du, dv):
            return sum(du[k] * dv[k] for k in range(dim))

        def fv(dv, x):
            return sum(dexact[k](x) * dv[k] for k in range(dim))

        def ff(v, x):
            return sum(dexact[k](x) ** 2 for k in range(dim)) * v
        M = self.iasm(uv)
        f = self.iasm(fv)
        uu = float(np.sum(self._unit_assembler().iasm(ff, intorder=intorder)))
        return np.sqrt(uu + np.dot(uh, M.dot(uh)) - 2. * np.dot(uh, f))

    def fasm(self, form, find=None, interior=False, intorder=None,
             normals=True, interp=None):
        """Assemble a bilinear or linear form over facets
        (spfem/assembly.py:1008-1183): boundary facets by default, interior
        facets (two-sided ``u1, u2, v1, v2, du1, ...`` forms) with
        ``interior=True``."""
        if find is None:
            find = (self.mesh.interior_facets() if interior
                    else self.mesh.boundary_facets())
        plan = self.plan_fasm(form, interior=interior, intorder=intorder,
                              normals=normals, interp=interp, find=find)
        return self._get_engine().run_facet(plan, np.asarray(find), interior,
                                            interp)
