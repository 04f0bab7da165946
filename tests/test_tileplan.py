# -*- coding: utf-8 -*-
"""Tile plan of the fused kernel checked on the CPU: a NumPy emulation of the
kernel's phase B over the plan must reproduce the reference result and write
every CSR value exactly once."""
import numpy as np
import pytest
import torch

import cases
import util
import hostemu
from spfem_b200.tileplan import (build_tile_plan, emulate, check_vertex_table,
                                 fill_coordinates, bank_conflicts)


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


def _pattern_numpy(rowdof, coldof, ncols):
    """Host restatement of spfem_pattern_sort/fill (stable sort by (row, col))."""
    nv, n = rowdof.shape
    nu = coldof.shape[0]
    ids = np.arange(nu * nv * n)
    blk, k = ids // n, ids % n
    j, i = blk // nv, blk % nv
    keys = rowdof[i, k].astype(np.int64) * ncols + coldof[j, k]
    order = np.argsort(keys, kind="stable")
    ks = keys[order]
    head = np.ones(len(ks), dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    cptr = np.concatenate((np.nonzero(head)[0], [len(ks)]))
    return ks[head], cptr, order


@pytest.mark.parametrize("name,E", [("trip1_lap", 64), ("trip1_lap_r6", 1024),
                                    ("trip2_lap", 32), ("stokes_b", 16),
                                    ("vecp2_elasticity", 50), ("tetp2_lap", 24)])
def test_tile_plan_reproduces_assembly(name, E):
    case = cases.BY_NAME[name]
    asm = util.assembler_for(case)
    plan = asm.plan_iasm(case.form)
    perm = _morton(asm.mesh)
    local = hostemu.run_interior(asm, plan)[:, perm]
    tv, tu = asm.dofnum_v.t_dof[:, perm], asm.dofnum_u.t_dof[:, perm]
    n = tv.shape[1]
    ukeys, cptr, contrib = _pattern_numpy(tv, tu, asm.dofnum_u.N)
    nnz = len(ukeys)
    rows = ukeys // asm.dofnum_u.N
    indptr = np.searchsorted(rows, np.arange(asm.dofnum_v.N + 1))
    vt = np.ascontiguousarray(asm.mesh.t[:, perm], dtype=np.int32)
    tp = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                         torch.from_numpy(contrib), n, plan.spec.n_entries, E,
                         torch.from_numpy(vt), dim=asm.mesh.p.shape[0],
                         alias=plan.spec.alias, schedule="full")
    fill_coordinates(tp, torch.from_numpy(np.ascontiguousarray(asm.mesh.p)))
    assert check_vertex_table(tp, vt, asm.mesh.p)
    # shared-memory image of the tile kernel: one row per distinct local value
    uniq = np.zeros((plan.spec.n_unique, local.shape[1]))
    uniq[plan.spec.alias, :] = local
    assert np.array_equal(uniq[plan.spec.alias, :], local)   # aliases really are equal
    values, written = emulate(tp, uniq, nnz)
    assert np.all(written == 1)
    ref = util.golden(case)
    assert nnz == ref.nnz
    scale = np.abs(ref.data).max()
    assert np.max(np.abs(values - ref.data)) <= 1e-12 * scale
    assert 1.0 <= tp["halo_factor"] < 3.0
    # the scheduled contribution lists are bank-conflict free
    assert tp["scheduled_rows"] > 0.9    # very long rows fall back to their CSR order
    if tp["scheduled_rows"] > 0.99:
        assert bank_conflicts(tp) < 1.05
    assert tp["list_growth"] < 2.5
    tp0 = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                          torch.from_numpy(contrib), n, plan.spec.n_entries, E,
                          torch.from_numpy(vt), dim=asm.mesh.p.shape[0],
                          alias=plan.spec.alias, schedule=False)
    v0, w0 = emulate(tp0, uniq, nnz)
    assert np.all(w0 == 1) and np.max(np.abs(v0 - ref.data)) <= 1e-12 * scale
    assert bank_conflicts(tp0) > 1.5
    # default: contributions re-ordered inside their slots only (no list growth)
    tp1 = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                          torch.from_numpy(contrib), n, plan.spec.n_entries, E,
                          torch.from_numpy(vt), dim=asm.mesh.p.shape[0],
                          alias=plan.spec.alias, schedule="inslot")
    v1, w1 = emulate(tp1, uniq, nnz)
    assert np.all(w1 == 1) and np.max(np.abs(v1 - ref.data)) <= 1e-12 * scale
    assert tp1["list_growth"] < 1.06
    # slot-parallel section B (one thread per CSR slot)
    tp2 = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                          torch.from_numpy(contrib), n, plan.spec.n_entries, E,
                          torch.from_numpy(vt), dim=asm.mesh.p.shape[0],
                          alias=plan.spec.alias, mode="slots")
    v2, w2 = emulate(tp2, uniq, nnz)
    assert np.all(w2 == 1) and np.max(np.abs(v2 - ref.data)) <= 1e-12 * scale
    assert tp2["halo_factor"] == tp1["halo_factor"]
    # ... with packed storage of the halo elements' values
    tp3 = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                          torch.from_numpy(contrib), n, plan.spec.n_entries, E,
                          torch.from_numpy(vt), dim=asm.mesh.p.shape[0],
                          alias=plan.spec.alias, mode="compact")
    v3, w3 = emulate(tp3, uniq, nnz)
    assert np.all(w3 == 1) and np.max(np.abs(v3 - ref.data)) <= 1e-12 * scale
    assert tp3["max_store"] < plan.spec.n_unique * tp3["ldl"]
    print(name, "dense", plan.spec.n_unique * tp3["ldl"], "compact", tp3["max_store"])
    # ... and the slots of a tile ordered by their number of contributions
    for md, srt in (("slots", True), ("compact", True), ("slots", "bucket")):
        tp4 = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                              torch.from_numpy(contrib), n, plan.spec.n_entries, E,
                              torch.from_numpy(vt), dim=asm.mesh.p.shape[0],
                              alias=plan.spec.alias, mode=md, sort_slots=srt)
        v4, w4 = emulate(tp4, uniq, nnz)
        assert np.all(w4 == 1) and np.max(np.abs(v4 - ref.data)) <= 1e-12 * scale


@pytest.mark.parametrize("ename", ["TriP1", "TriP2"])
def test_tile_plan_on_irregular_mesh(ename):
    """A mesh with a high-valence vertex (wheel of 40 triangles, refined and
    perturbed): rows of very different lengths, slots with 1 ... 40
    contributions, through every slot-parallel variant of the plan."""
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    from test_fanplan import _wheel_mesh
    mesh = _wheel_mesh(40, 1)
    asm = AssemblerElement(mesh, util.element(ename, 1))
    plan = asm.plan_iasm(cases.lap_mass if hasattr(cases, "lap_mass") else cases.lap2)
    ref = hostemu.iasm(asm, cases.lap_mass if hasattr(cases, "lap_mass") else cases.lap2)
    perm = _morton(mesh)
    local = hostemu.run_interior(asm, plan)[:, perm]
    tv, tu = asm.dofnum_v.t_dof[:, perm], asm.dofnum_u.t_dof[:, perm]
    n = tv.shape[1]
    ukeys, cptr, contrib = _pattern_numpy(tv, tu, asm.dofnum_u.N)
    nnz = len(ukeys)
    indptr = np.searchsorted(ukeys // asm.dofnum_u.N, np.arange(asm.dofnum_v.N + 1))
    vt = np.ascontiguousarray(mesh.t[:, perm], dtype=np.int32)
    uniq = np.zeros((plan.spec.n_unique, n))
    uniq[plan.spec.alias, :] = local
    scale = np.abs(ref.data).max()
    for md, srt in (("slots", False), ("compact", False), ("slots", True), ("compact", True),
                    ("slots", "bucket"), ("rows", False)):
        tp = build_tile_plan(torch, torch.from_numpy(indptr), torch.from_numpy(cptr),
                             torch.from_numpy(contrib), n, plan.spec.n_entries, 32,
                             torch.from_numpy(vt), dim=2, alias=plan.spec.alias,
                             mode=md, sort_slots=srt)
        v, w = emulate(tp, uniq, nnz)
        assert np.all(w == 1), (md, srt)
        assert np.max(np.abs(v - ref.data)) <= 1e-12 * scale, (md, srt)
