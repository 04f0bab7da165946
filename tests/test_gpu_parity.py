# -*- coding: utf-8 -*-
"""GPU parity tests proper: the product API (``AssemblerElement.iasm/fasm`` ->
C-ABI -> generated sm_100a kernels) against the reference's golden outputs,
the oracle, and size-independent properties at larger sizes."""
import numpy as np
import pytest

import cases
import util

pytestmark = pytest.mark.gpu


def _run(case):
    asm = util.assembler_for(case)
    kw = cases.case_inputs(case, asm.mesh, asm.dofnum_u.N)
    fn = asm.iasm if case.kind == "iasm" else asm.fasm
    return fn(case.form, **kw)


@pytest.mark.parametrize("case", cases.CASES, ids=lambda c: c.name)
def test_cuda_matches_reference_golden(case):
    util.check(_run(case), util.golden(case))


@pytest.mark.parametrize("name", ["trip1_lap_r6", "vecp2_elasticity", "tetp2_lap",
                                  "f_trip2_nitsche", "trip1_load_sin_h"])
def test_cuda_matches_oracle(name):
    case = cases.BY_NAME[name]
    util.check(_run(case), util.oracle_run(case))


def test_native_library_loaded():
    """The CUDA path must be the one that ran: the C-ABI library is mapped
    into this process and at least one generated module is resident."""
    from spfem_b200 import engine, _lib
    _run(cases.BY_NAME["trip1_lap"])
    assert _lib._lib is not None
    assert len(engine._MODULES) > 0
    with open("/proc/self/maps") as fh:
        assert "libspfem_b200.so" in fh.read()


def test_bitwise_reproducible():
    """No floating point atomics: two runs give identical bits."""
    case = cases.BY_NAME["vecp2_elasticity"]
    A, B = _run(case), _run(case)
    assert np.array_equal(A.data.view(np.int64), B.data.view(np.int64))
    case = cases.BY_NAME["trip1_load_sin_h"]
    a, b = _run(case), _run(case)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))


def test_readme_example_config1():
    """BASELINE config 1 (README.md:30-48): refine(6) P1 Poisson, direct solve;
    max(x) of the patched reference is 0.07365718549079245."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTri()
    m.refine(6)
    a = AssemblerElement(m, sp.ElementTriP1())
    A = a.iasm(lambda du, dv: du[0] * dv[0] + du[1] * dv[1])
    b = a.iasm(lambda v: 1.0 * v)
    assert A.shape == (4225, 4225) and A.nnz == 29057
    assert A.indices.dtype == np.int32 and A.has_sorted_indices
    x = sp.direct(A, b, I=m.interior_nodes())
    assert abs(np.max(x) - 0.07365718549079245) < 1e-12


def test_known_answer_test_asm_163():
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTri()
    m.refine(5)
    a = AssemblerElement(m, sp.ElementTriP1())
    A = a.iasm(cases.lap2)
    b = a.iasm(cases.load1)
    x = sp.direct(A, b, I=m.interior_nodes())
    assert abs(np.max(x) - 0.073614737354524146) < 1e-7


def test_large_mesh_properties():
    """refine(9) (524 288 triangles): properties that do not need the oracle --
    stiffness rows sum to zero, mass sums to the area, both symmetric, pattern
    nnz = nv + 2*ne, and subset assemblies add up to the whole
    (spfem/test_asm.py:274-310)."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTri()
    m.refine(9)
    a = AssemblerElement(m, sp.ElementTriP1())
    K = a.iasm(cases.lap2)
    M = a.iasm(cases.mass)
    nv, ne = m.p.shape[1], m.facets.shape[1]
    assert K.nnz == nv + 2 * ne == M.nnz
    assert np.max(np.abs(K.dot(np.ones(nv)))) < 1e-10
    assert abs(M.sum() - 1.0) < 1e-12
    assert abs(K - K.T).max() < 1e-12 and abs(M - M.T).max() < 1e-15
    nt = m.t.shape[1]
    K1 = a.iasm(cases.lap2, tind=np.arange(0, nt // 2))
    K2 = a.iasm(cases.lap2, tind=np.arange(nt // 2, nt))
    assert abs((K1 + K2) - K).max() < 1e-10
    f = a.iasm(cases.load1)
    assert abs(f.sum() - 1.0) < 1e-12


def test_device_resident_result():
    case = cases.BY_NAME["trip2_lap"]
    asm = util.assembler_for(case)
    d = asm.iasm_device(case.form)
    assert d.values.is_cuda and d.nnz == util.golden(case).nnz
    util.check(d.to_scipy(), util.golden(case))


@pytest.mark.parametrize("name", ["trip1_lap_r6", "trip2_lap", "vecp2_elasticity",
                                  "stokes_b", "tetp2_lap", "q2_lap",
                                  "trip1_interp_mass", "trip1_mass_x2"])
@pytest.mark.parametrize("mode", ["unfused", "dense", "dense_sorted", "bucket_sorted",
                                  "packed_values", "packed_sorted", "scalar_slots",
                                  "batched"])
def test_both_assembly_paths(name, mode, monkeypatch):
    """The two-kernel path (local matrices in HBM + gather reduction) and the
    fused tile kernel with many small tiles (exercises halos and row ownership;
    every variant of the native plan: dense / packed local values, slots in CSR
    order or ordered by contribution count, block or scalar slots for vector
    elements, several batches = several launches) reproduce the reference."""
    if mode == "unfused":
        monkeypatch.setenv("SPFEM_FUSED", "0")
    else:
        monkeypatch.setenv("SPFEM_TILE_E", "32")
        monkeypatch.setenv("SPFEM_TILE_SORT", "0")
        monkeypatch.setenv("SPFEM_TILE_MODE", "dense")
        if mode == "dense_sorted":
            monkeypatch.setenv("SPFEM_TILE_SORT", "1")
        elif mode == "bucket_sorted":
            monkeypatch.setenv("SPFEM_TILE_SORT", "bucket")
        elif mode == "packed_values":               # only the values that are read
            monkeypatch.setenv("SPFEM_TILE_MODE", "compact")
        elif mode == "packed_sorted":
            monkeypatch.setenv("SPFEM_TILE_MODE", "compact")
            monkeypatch.setenv("SPFEM_TILE_SORT", "1")
        elif mode == "scalar_slots":                # vector elements without block slots
            monkeypatch.setenv("SPFEM_TILE_BLOCKS", "0")
        elif mode == "batched":
            monkeypatch.setenv("SPFEM_TILE_BATCH", "2000")
        monkeypatch.setenv("SPFEM_FUSED", "force")   # bypass the "is it worth it" heuristic
        monkeypatch.setenv("SPFEM_FAN", "0")         # (the tile kernel, not the vertex-fan kernel)
    case = cases.BY_NAME[name]
    asm = util.assembler_for(case)
    kw = cases.case_inputs(case, asm.mesh, asm.dofnum_u.N)
    prep = asm.prepare_iasm(case.form, **kw)
    if mode == "unfused":
        assert prep.n_launches == 2
    else:
        assert type(prep).__name__ == "TiledPrepared"
        assert prep.n_launches == len(prep.tp["batches"])
        if mode == "batched" and prep.tp["n_contrib"] > 4000:
            assert prep.n_launches > 1
    prep.launch()
    util.check(prep.result(), util.golden(case))
    prep.launch()                                   # bitwise reproducible
    v1 = prep.values.clone()
    prep.launch()
    import torch
    assert torch.equal(v1, prep.values)


@pytest.mark.parametrize("name", ["trip1_lap", "trip1_mass", "trip1_lap_r6"])
@pytest.mark.parametrize("rows", ["256", "32"])
def test_vertex_fan_kernel(name, rows, monkeypatch):
    """Scalar P1 on triangles with element-constant coefficients runs the
    vertex-fan kernel (one thread per CSR row, no local matrices in shared
    memory)."""
    monkeypatch.setenv("SPFEM_FAN_ROWS", rows)
    case = cases.BY_NAME[name]
    asm = util.assembler_for(case)
    prep = asm.prepare_iasm(case.form)
    assert type(prep).__name__ == "FanPrepared" and prep.n_launches == 1
    prep.launch()
    A = prep.result()
    util.check(A, util.golden(case))
    prep.launch()
    assert np.array_equal(A.data.view(np.int64), prep.result().data.view(np.int64))


def test_vertex_fan_irregular_mesh_and_refresh():
    """Irregular valences; coordinates changed in place + upload_mesh()."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    import asm_oracle
    m = sp.MeshTri(initmesh="symmetric")
    m.refine(4)
    rng = np.random.default_rng(1)
    I = m.interior_nodes()
    m.p[:, I] += 0.005 * rng.standard_normal((2, len(I)))
    a = AssemblerElement(m, sp.ElementTriP1())
    form = lambda u, v, du, dv, h: h * (du[0] * dv[0] + 2.0 * du[1] * dv[1]) + u * v + du[0] * v
    prep = a.prepare_iasm(form)
    assert type(prep).__name__ == "FanPrepared"
    prep.launch()
    cases.compare_csr(prep.result(), asm_oracle.iasm(m, "TriP1", form))
    m.p[:, I] += 0.002 * rng.standard_normal((2, len(I)))
    a._get_engine().upload_mesh()
    prep.launch()
    cases.compare_csr(prep.result(), asm_oracle.iasm(m, "TriP1", form))


def test_vertex_fan_wide_records():
    """Valence-10 vertex: the 12-neighbour instantiation / 48-byte records."""
    import asm_oracle
    from test_fanplan import _wheel_mesh
    from spfem_b200.assembly import AssemblerElement
    import spfem_b200 as sp
    m = _wheel_mesh(10, 3)
    a = AssemblerElement(m, sp.ElementTriP1())
    prep = a.prepare_iasm(cases.lap2)
    assert type(prep).__name__ == "FanPrepared" and prep.fp["max_neighbours"] == 10
    prep.launch()
    cases.compare_csr(prep.result(), asm_oracle.iasm(m, "TriP1", cases.lap2))


def test_empty_and_ragged_inputs():
    """Empty element / facet subsets give all-zero results of the right shape;
    boolean masks and unsorted, repeated index lists behave like the reference
    (every listed element is integrated once per occurrence)."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    import asm_oracle
    m = sp.MeshTri()
    m.refine(3)
    a = AssemblerElement(m, sp.ElementTriP1())
    A = a.iasm(cases.lap2, tind=np.zeros(0, dtype=np.int64))
    assert A.shape == (81, 81) and A.nnz == 0 and A.indptr.dtype == np.int32
    f = a.iasm(cases.load1, tind=np.zeros(0, dtype=np.int64))
    assert f.shape == (81,) and not f.any()
    B = a.fasm(cases.bnd_mass, find=np.zeros(0, dtype=np.int64))
    assert B.shape == (81, 81) and B.nnz == 0
    tind = np.array([5, 3, 3, 100, 7])
    cases.compare_csr(a.iasm(cases.mass, tind=tind),
                      asm_oracle.iasm(m, "TriP1", cases.mass, tind=tind))
    mask = np.zeros(m.t.shape[1], dtype=bool)
    mask[::5] = True
    cases.compare_csr(a.iasm(cases.mass, tind=mask),
                      asm_oracle.iasm(m, "TriP1", cases.mass, tind=np.nonzero(mask)[0]))


def test_intorder_variants_and_quadrature_limits():
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    import asm_oracle
    m = sp.MeshTri()
    m.refine(2)
    a = AssemblerElement(m, sp.ElementTriP2())
    for order in (1, 2, 7, 12, 19):
        cases.compare_csr(a.iasm(cases.mass_x2, intorder=order),
                          asm_oracle.iasm(m, "TriP2", cases.mass_x2, intorder=order))
    with pytest.raises(NotImplementedError):
        a.iasm(cases.mass, intorder=20)
    mt = sp.MeshTet()
    mt.refine(1)
    at = AssemblerElement(mt, sp.ElementTetP1())
    with pytest.raises(NotImplementedError):
        at.iasm(cases.mass, intorder=5)


def test_full_size_config2_properties(monkeypatch):
    """BASELINE configs[1] at full size (MeshTri.refine(11), 8 388 608
    triangles): size-independent properties, and two independent kernels (the
    vertex-fan kernel and the two-kernel route with its radix-sorted
    contribution lists) agree on every one of the 29 372 417 CSR values."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTri()
    m.refine(11)
    nv, ne, nt = m.p.shape[1], m.facets.shape[1], m.t.shape[1]
    assert nt == 8388608 and nv == 4198401
    a = AssemblerElement(m, sp.ElementTriP1())
    prep = a.prepare_iasm(cases.lap2)
    assert type(prep).__name__ == "FanPrepared"
    prep.launch()
    K = prep.result()
    M = a.iasm(cases.mass)
    assert K.nnz == nv + 2 * ne == 29372417 and K.indices.dtype == np.int32
    assert np.array_equal(K.indptr, M.indptr) and np.array_equal(K.indices, M.indices)
    assert np.max(np.abs(K.dot(np.ones(nv)))) < 1e-9          # constants are in the kernel
    assert abs(M.sum() - 1.0) < 1e-11                          # area of the unit square
    assert abs(M.dot(m.p[0]).sum() - 0.5) < 1e-11              # first moment
    d = K - K.T
    assert abs(d).max() < 1e-12
    monkeypatch.setenv("SPFEM_FUSED", "0")
    a2 = AssemblerElement(m, sp.ElementTriP1())
    K2 = a2.iasm(cases.lap2)
    assert np.array_equal(K.indptr, K2.indptr) and np.array_equal(K.indices, K2.indices)
    scale = np.abs(K2.data).max()
    assert np.max(np.abs(K.data - K2.data)) <= 1e-12 * scale


@pytest.mark.parametrize("mname,ename,ref,dim", [("MeshTri", "TriP1", 4, 2), ("MeshTri", "TriP2", 3, 2),
                                                ("MeshTet", "TetP1", 2, 3), ("MeshQuad", "Q1", 3, 2)])
def test_error_norms_match_reference(mname, ename, ref, dim):
    """L2error / H1error (spfem/assembly.py:1353-1485) through the device path
    against values computed by the reference itself (tests/golden/error_norms.json)."""
    import json
    import os
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    gold = json.load(open(os.path.join(util.GOLDEN, "error_norms.json")))[mname + "_" + ename]
    m = getattr(sp, mname)()
    m.refine(ref)
    a = AssemblerElement(m, util.element(ename, 1))
    ex, dex = cases.exact_fun(dim)
    uh = cases.nodal_values(m, a, ex)
    assert abs(a.L2error(uh, ex) - gold["L2"]) <= 1e-9 * max(gold["L2"], 1e-3)
    assert abs(a.H1error(uh, dex) - gold["H1"]) <= 1e-9 * max(gold["H1"], 1e-3)


def test_rt0_mixed_poisson_convergence():
    """The reference's ``RT0Test`` (spfem/test_module.py:18-77): mixed Poisson
    with Raviart-Thomas fluxes and piecewise constants, L2 convergence rate of
    the scalar unknown between 0.95 and 1.15."""
    import scipy.sparse as spsp
    from scipy.sparse.linalg import spsolve
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    mesh = sp.MeshTri()
    mesh.refine(2)

    def fv(v, x):
        return 2 * np.pi ** 2 * np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]) * v

    def exact(x):
        return np.sin(np.pi * x[0]) * np.sin(np.pi * x[1])
    hs, errs = [], []
    for _ in range(1, 4):
        mesh.refine()
        a = AssemblerElement(mesh, sp.element.ElementTriRT0())
        b = AssemblerElement(mesh, sp.element.ElementTriRT0(), sp.element.ElementP0())
        c = AssemblerElement(mesh, sp.element.ElementP0())
        A = a.iasm(cases.rt0_sigtau)
        B = b.iasm(cases.rt0_divsigv)
        C = c.iasm(lambda u, v: u * v)
        f = c.iasm(fv)
        K = spsp.vstack((spsp.hstack((-A, -B.T)), spsp.hstack((-B, 0 * C)))).tocsr()
        u = spsolve(K, np.hstack((np.zeros(A.shape[0]), f)))
        Iu = np.arange(C.shape[0], dtype=np.int64) + A.shape[0]
        hs.append(mesh.param())
        errs.append(c.L2error(u[Iu], exact))
    rate = np.polyfit(np.log10(hs), np.log10(errs), 1)[0]
    assert 0.95 <= rate <= 1.15


@pytest.mark.parametrize("ename", ["TriP1", "TriP2"])
def test_tile_kernel_irregular_mesh(ename):
    """A high-valence vertex (wheel of 40 triangles, refined): the vertex-fan
    plan declines (> 12 ring neighbours), the element-tile kernel takes over;
    rows of very different lengths and slots with 1 ... 40 contributions."""
    import asm_oracle
    from test_fanplan import _wheel_mesh
    from spfem_b200.assembly import AssemblerElement
    mesh = _wheel_mesh(40, 2)
    a = AssemblerElement(mesh, util.element(ename, 1))
    prep = a.prepare_iasm(cases.lap2)
    assert type(prep).__name__ == "TiledPrepared"
    prep.launch()
    cases.compare_csr(prep.result(), asm_oracle.iasm(mesh, ename, cases.lap2))
    cases.compare_vec(a.iasm(cases.load1), asm_oracle.iasm(mesh, ename, cases.load1))
