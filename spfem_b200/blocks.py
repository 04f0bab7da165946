# -*- coding: utf-8 -*-
"""Block systems on the device.

The reference builds saddle-point / multi-field systems on the host:
``spfem.utils.stack`` (``scipy.sparse.vstack/hstack``, spfem/utils.py:12-20),
Dirichlet elimination by slicing ``A[I].T[I].T`` and a direct solve
(``spfem.utils.direct``, :52-95).  With assembly on the GPU that costs one
device-to-host copy per block and a host merge.  Here the blocks produced by
``AssemblerElement.iasm_device`` (:class:`engine.DeviceCSR`) stay where they
are:

* :func:`stack` -- ``[[A, B.T], [B, None]]``-style block stacking into one CSR,
* :func:`condense` -- the sub-system on an index set ``I`` with the Dirichlet
  values moved to the right-hand side (``direct``'s elimination),
* :func:`cg` -- Jacobi-preconditioned conjugate gradients on the device for
  the symmetric positive definite systems (Poisson, elasticity); the sparse
  matrix-vector product is cuSPARSE through ``torch.sparse_csr_tensor``.

Everything is index arithmetic on tensors (set-up plumbing, like the plans);
the functions also run on CPU tensors, which is how ``tests/test_blocks.py``
checks them against SciPy without a GPU.  The matrices are :class:`CSR`
triples ``(indptr, indices, values, shape)`` with sorted indices and the
explicit zeros of the assembly kept.
"""
import numpy as np


class CSR(object):
    """Plain CSR triple on one device (int64 ``indptr`` / ``indices``)."""

    def __init__(self, indptr, indices, values, shape):
        self.indptr, self.indices, self.values = indptr, indices, values
        self.shape = (int(shape[0]), int(shape[1]))

    @property
    def nnz(self):
        return int(self.indices.numel())

    @property
    def device(self):
        return self.values.device

    @property
    def T(self):
        return _Transposed(self)

    def coo(self, torch):
        counts = self.indptr[1:] - self.indptr[:-1]
        rows = torch.repeat_interleave(
            torch.arange(self.shape[0], device=self.device), counts)
        return rows, self.indices, self.values

    def to_torch(self, torch):
        return torch.sparse_csr_tensor(self.indptr, self.indices, self.values,
                                       size=self.shape, check_invariants=False)

    def to_scipy(self):
        from scipy.sparse import csr_matrix
        return csr_matrix((self.values.cpu().numpy(), self.indices.cpu().numpy(),
                           self.indptr.cpu().numpy()), shape=self.shape)

    def matvec(self, torch, x):
        return torch.mv(self.to_torch(torch), x)

    def diagonal(self, torch):
        rows, cols, vals = self.coo(torch)
        d = torch.zeros(self.shape[0], dtype=vals.dtype, device=self.device)
        on = rows == cols
        d.index_add_(0, rows[on], vals[on])
        return d


class _Transposed(object):
    """Marker for a transposed block in :func:`stack`."""

    def __init__(self, mat):
        self.mat = mat


def as_csr(torch, A):
    """:class:`CSR` view of an ``engine.DeviceCSR`` / SciPy matrix / CSR."""
    if isinstance(A, CSR):
        return A
    if hasattr(A, "pattern"):                      # engine.DeviceCSR
        p = A.pattern
        return CSR(p.indptr.to(torch.int64), p.indices.to(torch.int64), A.values, A.shape)
    if hasattr(A, "indptr"):                       # scipy.sparse (tests)
        A = A.tocsr()
        return CSR(torch.from_numpy(A.indptr.astype(np.int64)),
                   torch.from_numpy(A.indices.astype(np.int64)),
                   torch.from_numpy(np.asarray(A.data, dtype=np.float64)), A.shape)
    raise TypeError("cannot interpret %r as a CSR matrix" % (type(A),))


def _from_coo(torch, rows, cols, vals, shape):
    """CSR with sorted indices from COO triplets without duplicates."""
    nrows, ncols = shape
    order = torch.argsort(rows * ncols + cols)
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = torch.zeros(nrows + 1, dtype=torch.int64, device=vals.device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=nrows), 0)
    return CSR(indptr, cols.contiguous(), vals.contiguous(), shape)


def transpose(torch, A):
    A = as_csr(torch, A)
    rows, cols, vals = A.coo(torch)
    return _from_coo(torch, cols, rows, vals, (A.shape[1], A.shape[0]))


def stack(torch, blocks):
    """Block matrix from a 2-D list of blocks (``spfem.utils.stack``):
    entries are matrices, ``M.T`` of a :class:`CSR`, ``("T", M)`` for the
    transpose of any accepted matrix, or ``None`` for a zero block."""
    mats = []
    for brow in blocks:
        out = []
        for blk in brow:
            if blk is None:
                out.append(None)
            elif isinstance(blk, _Transposed):
                out.append(transpose(torch, blk.mat))
            elif isinstance(blk, tuple) and len(blk) == 2 and blk[0] == "T":
                out.append(transpose(torch, blk[1]))
            else:
                out.append(as_csr(torch, blk))
        mats.append(out)
    nbr, nbc = len(mats), len(mats[0])
    heights = [next((m.shape[0] for m in row if m is not None), None) for row in mats]
    widths = [next((mats[r][c].shape[1] for r in range(nbr) if mats[r][c] is not None), None)
              for c in range(nbc)]
    if None in heights or None in widths:
        raise ValueError("a block row / column consists of zero blocks only")
    r0 = np.concatenate(([0], np.cumsum(heights)))
    c0 = np.concatenate(([0], np.cumsum(widths)))
    R, Cc, V = [], [], []
    for r in range(nbr):
        for c in range(nbc):
            m = mats[r][c]
            if m is None:
                continue
            if m.shape != (heights[r], widths[c]):
                raise ValueError("block (%d, %d) has shape %s, expected %s"
                                 % (r, c, m.shape, (heights[r], widths[c])))
            rows, cols, vals = m.coo(torch)
            R.append(rows + int(r0[r]))
            Cc.append(cols + int(c0[c]))
            V.append(vals)
    return _from_coo(torch, torch.cat(R), torch.cat(Cc), torch.cat(V),
                     (int(r0[-1]), int(c0[-1])))


def condense(torch, A, b, I, x=None):
    """``(A_II, b_I - A_ID x_D)`` for the index set ``I`` (tensor or array):
    the Dirichlet elimination of ``spfem.utils.direct`` (:84-93).  ``x`` holds
    the Dirichlet values (zero if omitted)."""
    A = as_csr(torch, A)
    dev = A.device
    I = torch.as_tensor(np.asarray(I) if not torch.is_tensor(I) else I,
                        dtype=torch.int64, device=dev)
    n = A.shape[0]
    new = torch.full((n,), -1, dtype=torch.int64, device=dev)
    new[I] = torch.arange(I.numel(), device=dev)
    rows, cols, vals = A.coo(torch)
    keep_r = new[rows] >= 0
    inner = keep_r & (new[cols] >= 0)
    AII = _from_coo(torch, new[rows[inner]], new[cols[inner]], vals[inner],
                    (int(I.numel()), int(I.numel())))
    bI = b[I].clone()
    if x is not None:
        outer = keep_r & (new[cols] < 0)
        bI.index_add_(0, new[rows[outer]], -vals[outer] * x[cols[outer]])
    return AII, bI, I


def cg(torch, A, b, tol=1e-10, maxiter=None, x0=None):
    """Jacobi-preconditioned conjugate gradients for a symmetric positive
    definite :class:`CSR` matrix.  Returns ``(x, iterations, relative residual)``."""
    A = as_csr(torch, A)
    M = A.to_torch(torch)
    dinv = 1.0 / A.diagonal(torch)
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b - torch.mv(M, x)
    z = dinv * r
    p = z.clone()
    rz = torch.dot(r, z)
    bnorm = torch.linalg.vector_norm(b)
    if float(bnorm) == 0.0:
        return x, 0, 0.0
    maxiter = 10 * b.numel() if maxiter is None else maxiter
    it, res = 0, float(torch.linalg.vector_norm(r) / bnorm)
    while res > tol and it < maxiter:
        Ap = torch.mv(M, p)
        alpha = rz / torch.dot(p, Ap)
        x += alpha * p
        r -= alpha * Ap
        z = dinv * r
        rz_new = torch.dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
        if it % 10 == 0 or it < 10:
            res = float(torch.linalg.vector_norm(r) / bnorm)
    return x, it, float(torch.linalg.vector_norm(r) / bnorm)


def solve_dirichlet(torch, A, b, I, x=None, tol=1e-10):
    """``spfem.utils.direct(A, b, x, I)`` with the elimination and a PCG solve on
    the device; returns the full solution vector (a tensor)."""
    A = as_csr(torch, A)
    AII, bI, I = condense(torch, A, b, I, x)
    xI, _, _ = cg(torch, AII, bI, tol=tol)
    out = torch.zeros(A.shape[0], dtype=b.dtype, device=b.device) if x is None else x.clone()
    out[I] = xI
    return out
