# -*- coding: utf-8 -*-
"""Native pattern + tile-plan builder (csrc/spfem_plan.cu) checked on the CPU:
``on_device = 0`` runs the same code on host pointers.  The CSR structure must
equal SciPy's ``coo_matrix(...).tocsr()`` bit for bit, and a NumPy emulation of
the kernel's phase B over the blobs must reproduce the reference result and
write every CSR value exactly once."""
import numpy as np
import pytest
import torch

import cases
import util
import hostemu
from spfem_b200.tileplan import (build_tile_plan, emulate, check_vertex_table,
                                 fill_coordinates, TileTooLarge)


def _morton(mesh):
    """Host Morton order of the element centroids (the device path uses
    spfem_morton_order)."""
    c = mesh.p[:, mesh.t].mean(axis=1)
    lo, hi = mesh.p.min(axis=1, keepdims=True), mesh.p.max(axis=1, keepdims=True)
    q = ((c - lo) / (hi - lo) * 1023).astype(np.int64)
    key = np.zeros(c.shape[1], dtype=np.int64)
    for b in range(10):
        for d in range(c.shape[0]):
            key |= ((q[d] >> b) & 1) << (b * c.shape[0] + d)
    return np.argsort(key, kind="stable")


def _ut(spec, bsv, bsu):
    """Distinct-value ids of the bsv x bsu block of every scalar entry (what
    codegen emits as SPFEM_UT)."""
    nbv, nbu = spec.n_rows // bsv, spec.n_cols // bsu
    ut = np.zeros((nbu * nbv, bsv * bsu), dtype=np.int64)
    for j in range(nbu):
        for i in range(nbv):
            for a in range(bsv):
                for b in range(bsu):
                    ut[j * nbv + i, a * bsu + b] = spec.alias[(bsu * j + b) * spec.n_rows
                                                              + (bsv * i + a)]
    return ut


def _setup(case_or_asm, form):
    asm = case_or_asm
    plan = asm.plan_iasm(form)
    perm = _morton(asm.mesh)
    local = hostemu.run_interior(asm, plan)[:, perm]
    tv = np.ascontiguousarray(asm.dofnum_v.t_dof[:, perm], dtype=np.int32)
    tu = np.ascontiguousarray(asm.dofnum_u.t_dof[:, perm], dtype=np.int32)
    vt = np.ascontiguousarray(asm.mesh.t[:, perm], dtype=np.int32)
    uniq = np.zeros((plan.spec.n_unique, local.shape[1]))
    uniq[plan.spec.alias, :] = local
    assert np.array_equal(uniq[plan.spec.alias, :], local)   # aliases really are equal
    return plan, tv, tu, vt, uniq


def _build(asm, plan, tv, tu, vt, E, **kw):
    spec = plan.spec
    tp = build_tile_plan(torch, torch.from_numpy(tv),
                         torch.from_numpy(tu) if plan.bilinear else None,
                         torch.from_numpy(vt), asm.dofnum_v.N,
                         asm.dofnum_u.N if plan.bilinear else 1, E,
                         spec.alias, spec.n_unique, asm.mesh.p.shape[0],
                         with_el=True, **kw)
    fill_coordinates(tp, torch.from_numpy(np.ascontiguousarray(asm.mesh.p)))
    return tp


VARIANTS = [dict(), dict(owner_rule=0), dict(owner_rule=0, compact=True, sort_slots=1), dict(sort_slots=1), dict(sort_slots=2), dict(compact=True),
            dict(compact=True, sort_slots=1), dict(batch_cap=3000),
            dict(batch_cap=1500, sort_slots=1), dict(split_factor=1.05),
            dict(split_factor=1.1, compact=True, sort_slots=1)]


@pytest.mark.parametrize("name,E", [("trip1_lap", 64), ("trip1_lap_r6", 1024),
                                    ("trip2_lap", 32), ("stokes_b", 16),
                                    ("vecp2_elasticity", 50), ("tetp2_lap", 24),
                                    ("vectetp1_mixed", 20)])
def test_native_plan_reproduces_assembly(name, E):
    case = cases.BY_NAME[name]
    asm = util.assembler_for(case)
    plan, tv, tu, vt, uniq = _setup(asm, case.form)
    ref = util.golden(case)
    scale = np.abs(ref.data).max()
    spec = plan.spec
    blocks = [(1, 1)]
    if spec.ncu * spec.ncv > 1:
        blocks.append((spec.ncv, spec.ncu))
    for bsv, bsu in blocks:
        for kw in VARIANTS:
            if kw.get("compact") and bsv * bsu > 1:
                continue
            tp = _build(asm, plan, tv, tu, vt, E, bsv=bsv, bsu=bsu, **kw)
            assert (tp["bsv"], tp["bsu"]) == (bsv, bsu)
            if "batch_cap" in kw and tp["n_contrib"] > 2 * kw["batch_cap"]:
                assert len(tp["batches"]) > 1
            assert check_vertex_table(tp, vt, asm.mesh.p)
            # the CSR structure is SciPy's, bit for bit
            assert tp["nnz"] == ref.nnz
            assert np.array_equal(tp["indptr"].numpy(), ref.indptr)
            assert np.array_equal(tp["indices"].numpy(), ref.indices)
            values, written = emulate(tp, uniq, _ut(spec, bsv, bsu))
            assert np.all(written == 1), (bsv, bsu, kw)
            assert np.max(np.abs(values - ref.data)) <= 1e-12 * scale, (bsv, bsu, kw)
            assert 1.0 <= tp["halo_factor"] < 4.0
            if kw.get("compact"):
                assert tp["max_store"] <= spec.n_unique * tp["ldl"]


def test_native_plan_int64_indices_and_linear_form():
    case = cases.BY_NAME["trip2_lap"]
    asm = util.assembler_for(case)
    plan, tv, tu, vt, uniq = _setup(asm, case.form)
    ref = util.golden(case)
    tp = _build(asm, plan, tv, tu, vt, 32, index_bytes=8)
    assert tp["indptr"].dtype == torch.int64 and tp["indices"].dtype == torch.int64
    assert np.array_equal(tp["indptr"].numpy(), ref.indptr)
    assert np.array_equal(tp["indices"].numpy(), ref.indices)
    # a load vector = one slot per row
    asm1 = util.assembler_for(cases.BY_NAME["trip1_lap"])
    plan, tv, tu, vt, uniq = _setup(asm1, cases.load1)
    tp = _build(asm1, plan, tv, tu, vt, 64)
    assert tp["indices"] is None
    values, written = emulate(tp, uniq)
    assert np.all(written == 1)
    ref = hostemu.iasm(asm1, cases.load1)
    assert np.max(np.abs(values - ref)) <= 1e-12 * np.abs(ref).max()


def test_native_plan_row_mask():
    """Rows that are not assembled here (another rank owns them) get no slots."""
    case = cases.BY_NAME["trip2_lap"]
    asm = util.assembler_for(case)
    plan, tv, tu, vt, uniq = _setup(asm, case.form)
    ref = util.golden(case)
    N = asm.dofnum_v.N
    mask = np.ones(N, dtype=np.uint8)
    mask[::3] = 0
    tp = _build(asm, plan, tv, tu, vt, 32, row_mask=torch.from_numpy(mask))
    indptr = tp["indptr"].numpy()
    lens = np.diff(indptr)
    assert np.all(lens[mask == 0] == 0)
    assert np.array_equal(lens[mask == 1], np.diff(ref.indptr)[mask == 1])
    values, written = emulate(tp, uniq)
    assert np.all(written == 1)
    keep = np.repeat(mask == 1, np.diff(ref.indptr))
    assert np.array_equal(tp["indices"].numpy(), ref.indices[keep])
    assert np.max(np.abs(values - ref.data[keep])) <= 1e-12 * np.abs(ref.data).max()


def test_native_plan_tile_too_large():
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    m = sp.MeshTet()
    m.refine(3)
    asm = AssemblerElement(m, sp.ElementTetP2())
    plan, tv, tu, vt, uniq = _setup(asm, cases.lap3)
    with pytest.raises(TileTooLarge):
        _build(asm, plan, tv, tu, vt, 4096)


@pytest.mark.parametrize("ename", ["TriP1", "TriP2"])
def test_native_plan_on_irregular_mesh(ename):
    """A mesh with a high-valence vertex (wheel of 40 triangles, refined and
    perturbed): rows of very different lengths, slots with 1 ... 40
    contributions, through every variant of the plan."""
    from spfem_b200.assembly import AssemblerElement
    from test_fanplan import _wheel_mesh
    mesh = _wheel_mesh(40, 1)
    asm = AssemblerElement(mesh, util.element(ename, 1))
    form = cases.lap_mass if hasattr(cases, "lap_mass") else cases.lap2
    plan, tv, tu, vt, uniq = _setup(asm, form)
    ref = hostemu.iasm(asm, form)
    scale = np.abs(ref.data).max()
    for kw in VARIANTS:
        tp = _build(asm, plan, tv, tu, vt, 32, **kw)
        assert np.array_equal(tp["indptr"].numpy(), ref.indptr)
        assert np.array_equal(tp["indices"].numpy(), ref.indices)
        v, w = emulate(tp, uniq)
        assert np.all(w == 1), kw
        assert np.max(np.abs(v - ref.data)) <= 1e-12 * scale, kw
