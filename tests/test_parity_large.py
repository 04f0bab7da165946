# -*- coding: utf-8 -*-
"""Parity at sizes where the tile / fan plans have thousands of tiles: the product
(C-ABI -> generated sm_100a kernels) against the ORACLE (not against itself) at
about half a million elements, the 64-bit index branches of the pattern and plan
builders, and one assembly whose local entries exceed 2^31-1 (SciPy's int64
rule, scipy/sparse/_coo.py)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _mesh(kind, n):
    import spfem_b200 as sp
    from spfem_b200 import meshgen
    if kind == "refine":
        m = sp.MeshTri()
        m.refine(n)
        return m
    if kind == "tri":
        p, t = meshgen.grid_tri(n)
        return sp.MeshTri(p, t, validate=False)
    p, t = meshgen.grid_tet(n)
    return sp.MeshTet(p, t, validate=False)


# name -> (mesh kind, size, oracle element, components, form, route the engine must take)
LARGE = {
    "trip1_refine9": ("refine", 9, "TriP1", 1, cases.lap2, "FanPrepared"),
    "trip2_grid512": ("tri", 512, "TriP2", 1, cases.lap2, "TiledPrepared"),
    "vecp2_grid512": ("tri", 512, "TriP2", 2, cases.elasticity, "TiledPrepared"),
    "tetp2_grid44": ("tet", 44, "TetP2", 1, cases.lap3, "TiledPrepared"),
}


@pytest.mark.parametrize("name", sorted(LARGE))
def test_product_matches_oracle_at_half_a_million_elements(name):
    """>= 5e5 elements: 4096+ fan tiles / 5000+ element tiles, halo elements, split
    tiles and several plan batches are all in play; the oracle is the NumPy
    restatement of spfem/assembly.py:919-1006 pinned to the reference's goldens."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    import asm_oracle
    import util
    kind, n, ename, ncomp, form, route = LARGE[name]
    m = _mesh(kind, n)
    assert m.t.shape[1] >= 500000
    a = AssemblerElement(m, util.element(ename, ncomp))
    prep = a.prepare_iasm(form)
    assert type(prep).__name__ == route
    prep.launch()
    A = prep.result()
    B = asm_oracle.iasm(m, ename, form, ncomp_u=ncomp, ncomp_v=ncomp)
    err = cases.compare_csr(A, B)          # identical pattern + 1e-12 relative per entry
    assert err <= 1e-12
    # and through the one-shot public call
    A2 = a.iasm(form)
    assert np.array_equal(A2.data.view(np.int64), A.data.view(np.int64))


def test_pattern_fill_int64_equals_int32():
    """``spfem_pattern_fill(index_bytes=8)`` (``k_fill_csr<long long>``) against the
    32-bit result of the same sorted workspace, through the C-ABI."""
    import ctypes as C
    import torch
    import spfem_b200 as sp
    from spfem_b200 import _lib
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTri()
    m.refine(5)
    a = AssemblerElement(m, sp.ElementTriP2(), sp.ElementTriP1())     # rectangular
    eng = a._get_engine()
    lib = eng.lib
    rowdof, coldof = eng.tdof_v.contiguous(), eng.tdof_u.contiguous()
    nloc_v, n = rowdof.shape
    nloc_u = coldof.shape[0]
    n_entries = n * nloc_v * nloc_u
    nrows, ncols = a.dofnum_v.N, a.dofnum_u.N
    out = {}
    for ib, tdt in ((4, torch.int32), (8, torch.int64)):
        wsb = lib.spfem_pattern_ws_bytes(n_entries)
        ws = torch.empty(wsb, dtype=torch.uint8, device=eng.device)
        nnz = C.c_longlong(0)
        _lib.check(lib.spfem_pattern_sort(eng._ptr(rowdof), n, eng._ptr(coldof), n, n, nloc_v,
                                          nloc_u, nrows, ncols, eng._ptr(ws), wsb, C.byref(nnz),
                                          eng._stream()))
        indptr = torch.empty(nrows + 1, dtype=tdt, device=eng.device)
        indices = torch.empty(nnz.value, dtype=tdt, device=eng.device)
        cptr = torch.empty(nnz.value + 1, dtype=torch.int32, device=eng.device)
        contrib = torch.empty(n_entries, dtype=torch.int32, device=eng.device)
        _lib.check(lib.spfem_pattern_fill(eng._ptr(ws), n_entries, nrows, ncols, nnz.value, ib,
                                          eng._ptr(indptr), eng._ptr(indices), eng._ptr(cptr),
                                          eng._ptr(contrib), None, eng._stream()))
        torch.cuda.synchronize()
        out[ib] = (indptr.cpu().numpy(), indices.cpu().numpy(), cptr.cpu().numpy(),
                   contrib.cpu().numpy())
    assert out[8][0].dtype == np.int64 and out[8][1].dtype == np.int64
    for k in range(4):
        assert np.array_equal(out[4][k], out[8][k].astype(out[4][k].dtype))
    ref = a.iasm(cases.mass)
    assert np.array_equal(out[8][0], ref.indptr) and np.array_equal(out[8][1], ref.indices)


@pytest.mark.parametrize("route", ["tile", "two-kernel", "chunked"])
def test_forced_int64_indices_give_the_int32_matrix(route, monkeypatch):
    """Every 64-bit index branch (native tile plan ``index_bytes = 8``, the sorted
    pattern of the two-kernel route, the device merge of the chunked assembly) on a
    mesh small enough to compare with the 32-bit run: same structure, same values."""
    import spfem_b200 as sp
    from spfem_b200 import engine
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTri()
    m.refine(4)

    def assemble():
        a = AssemblerElement(m, sp.ElementTriP2())
        if route == "chunked":
            monkeypatch.setenv("SPFEM_MAX_ENTRIES", str(36 * m.t.shape[1] // 3))
        if route == "two-kernel":
            return a.iasm(cases.lap2, tind=np.arange(m.t.shape[1]))
        return a.iasm(cases.lap2)
    A32 = assemble()
    assert A32.indices.dtype == np.int32
    monkeypatch.setattr(engine, "scipy_index_dtype", lambda *args: np.int64)
    A64 = assemble()
    assert A64.indices.dtype == np.int64 and A64.indptr.dtype == np.int64
    assert np.array_equal(A64.indptr, A32.indptr) and np.array_equal(A64.indices, A32.indices)
    assert np.max(np.abs(A64.data - A32.data)) <= 1e-14 * np.abs(A32.data).max()


def test_more_than_2_31_local_entries_int64_device_resident():
    """BASELINE config 3 at full size: vector-P2 elasticity on the 3162^2 grid has
    144 * 19 996 488 = 2.88e9 > 2^31-1 local entries, so SciPy's rule gives int64
    indices.  One fused launch per plan batch, result device resident; checked
    through size-independent properties: the CSR structure is consistent, the
    matrix is symmetric in its row sums and annihilates rigid translations."""
    import torch
    import spfem_b200 as sp
    from spfem_b200 import meshgen
    from spfem_b200.assembly import AssemblerElement
    if torch.cuda.get_device_properties(0).total_memory < 100 * 2 ** 30:
        pytest.skip("needs ~70 GB of device memory")
    p, t = meshgen.grid_tri(3162)
    m = sp.MeshTri(p, t, validate=False)
    a = AssemblerElement(m, sp.ElementH1Vec(sp.ElementTriP2()))
    assert 144 * m.t.shape[1] > 2 ** 31 - 1
    A = a.iasm_device(cases.elasticity)
    pat = A.pattern
    assert pat.indptr.dtype == torch.int64 and pat.indices.dtype == torch.int64
    N = a.dofnum_u.N
    assert A.shape == (N, N) and int(pat.indptr[-1]) == pat.nnz == A.values.numel()
    assert bool((pat.indptr[1:] >= pat.indptr[:-1]).all())
    assert int(pat.indices.min()) >= 0 and int(pat.indices.max()) == N - 1
    # rigid translations have zero strain: A @ (1, 0, 1, 0, ...) = 0 = A @ (0, 1, 0, 1, ...)
    rows = torch.repeat_interleave(torch.arange(N, device=A.values.device),
                                   pat.indptr[1:] - pat.indptr[:-1])
    scale = float(A.values.abs().max())
    for comp in (0, 1):
        sel = (pat.indices % 2) == comp
        y = torch.zeros(N, dtype=torch.float64, device=A.values.device)
        y.index_add_(0, rows[sel], A.values[sel])
        assert float(y.abs().max()) <= 1e-10 * scale
    # trace of the stiffness matrix is positive and every diagonal entry is stored
    diag = A.values[(pat.indices == rows)]
    assert diag.numel() == N and float(diag.min()) > 0.0


def test_config4_full_size_tetp2_energy_identities():
    """BASELINE config 4 at full size: ElementTetP2 Laplace on the Kuhn 203^3 grid
    (50 193 162 tetrahedra, ~2e9 nonzeros, 5e9 local entries => int64 indices).
    Size-independent properties of the assembled matrix, with the P2 interpolants of
    1, x and y^2 (P2 reproduces them exactly): A 1 = 0, x^T A x = |grad x|^2 |cube| = 1,
    (y^2)^T A (y^2) = int (2y)^2 = 4/3, x^T A y^2 = 0."""
    import torch
    import spfem_b200 as sp
    from spfem_b200 import meshgen
    from spfem_b200.assembly import AssemblerElement
    if torch.cuda.get_device_properties(0).total_memory < 150 * 2 ** 30:
        pytest.skip("needs ~120 GB of device memory")
    p, t = meshgen.grid_tet(203)
    m = sp.MeshTet(p, t, validate=False)
    assert m.t.shape[1] == 6 * 203 ** 3
    a = AssemblerElement(m, sp.ElementTetP2())
    A = a.iasm_device(cases.lap3)
    pat = A.pattern
    N = a.dofnum_u.N
    assert pat.indptr.dtype == torch.int64 and A.shape == (N, N)
    assert int(pat.indptr[-1]) == pat.nnz == A.values.numel()
    dev = A.values.device
    # coordinates of the DOFs: vertices, then edge midpoints (spfem/assembly.py:1499-1512)
    X = np.empty((3, N))
    X[:, a.dofnum_u.n_dof[0]] = m.p
    X[:, a.dofnum_u.e_dof[0]] = 0.5 * (m.p[:, m.edges[0]] + m.p[:, m.edges[1]])
    one = torch.ones(N, dtype=torch.float64, device=dev)
    ux = torch.from_numpy(X[0]).to(dev)
    uy2 = torch.from_numpy(X[1] ** 2).to(dev)
    lens = pat.indptr[1:] - pat.indptr[:-1]

    def matvec(u):                           # row blocks: bounded scratch memory
        y = torch.empty(N, dtype=torch.float64, device=dev)
        step = 1 << 22
        for r0 in range(0, N, step):
            r1 = min(N, r0 + step)
            k0, k1 = int(pat.indptr[r0]), int(pat.indptr[r1])
            rows = torch.repeat_interleave(torch.arange(r1 - r0, device=dev), lens[r0:r1])
            part = torch.zeros(r1 - r0, dtype=torch.float64, device=dev)
            part.index_add_(0, rows, A.values[k0:k1] * u[pat.indices[k0:k1]])
            y[r0:r1] = part
        return y
    scale = float(A.values.abs().max())
    assert float(matvec(one).abs().max()) <= 1e-10 * scale
    Ax = matvec(ux)
    assert abs(float(ux @ Ax) - 1.0) <= 1e-9
    Ay2 = matvec(uy2)
    assert abs(float(uy2 @ Ay2) - 4.0 / 3.0) <= 1e-9
    assert abs(float(ux @ Ay2)) <= 1e-9
    assert abs(float(uy2 @ Ax)) <= 1e-9


def test_config5_full_size_divergence_block_identities():
    """BASELINE config 5 at full size: the Taylor-Hood divergence block B (rows P1,
    columns vector-P2, ``(du[0][0]+du[1][1])*v``, learning/Example 1 - Stokes
    equations.ipynb cell 9) on the 3162^2 grid.  With the P2 interpolant of a velocity
    field U, (B U)_i = int div(U) phi_i: div (x, y) = 2 => sum_i (B U)_i = 2 |square| = 2 and
    B U = 2 M 1 >= 0; the rotation (-y, x) is divergence free => B U = 0."""
    import torch
    import spfem_b200 as sp
    from spfem_b200 import meshgen
    from spfem_b200.assembly import AssemblerElement
    if torch.cuda.get_device_properties(0).total_memory < 100 * 2 ** 30:
        pytest.skip("needs ~40 GB of device memory")
    p, t = meshgen.grid_tri(3162)
    m = sp.MeshTri(p, t, validate=False)
    a = AssemblerElement(m, sp.ElementH1Vec(sp.ElementTriP2()), sp.ElementTriP1())
    B = a.iasm_device(cases.stokes_b)
    pat = B.pattern
    Nv, Nu = a.dofnum_v.N, a.dofnum_u.N
    assert B.shape == (Nv, Nu) and Nv == m.p.shape[1]
    dev = B.values.device
    # DOF coordinates of the vector-P2 space: components interleaved at vertices, then at
    # facet (edge) midpoints (spfem/assembly.py:1499-1521, element.py:505-507)
    dn = a.dofnum_u
    X = np.empty((2, Nu))
    mid = 0.5 * (m.p[:, m.facets[0]] + m.p[:, m.facets[1]])
    for c in range(2):
        X[:, dn.n_dof[c]] = m.p
        X[:, dn.f_dof[c]] = mid
    comp = np.empty(Nu, dtype=np.int64)
    for c in range(2):
        comp[dn.n_dof[c]] = c
        comp[dn.f_dof[c]] = c
    radial = torch.from_numpy(np.where(comp == 0, X[0], X[1])).to(dev)
    rot = torch.from_numpy(np.where(comp == 0, -X[1], X[0])).to(dev)
    rows = torch.repeat_interleave(torch.arange(Nv, device=dev), pat.indptr[1:] - pat.indptr[:-1])

    def matvec(u):
        y = torch.zeros(Nv, dtype=torch.float64, device=dev)
        y.index_add_(0, rows, B.values * u[pat.indices])
        return y
    y = matvec(radial)
    assert abs(float(y.sum()) - 2.0) <= 1e-10
    assert float(y.min()) > 0.0
    scale = float(y.max())
    assert float(matvec(rot).abs().max()) <= 1e-9 * scale
