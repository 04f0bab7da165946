# -*- coding: utf-8 -*-
"""Vertex-fan kernel on the CPU: ring construction, plan and the generated
``_fan_host`` loop against the reference's golden matrices."""
import numpy as np
import pytest

import spfem_b200 as sp
from spfem_b200.assembly import AssemblerElement
from fanplan_ref import vertex_rings, build_fan_plan, FanPlanError
import cases
import util
import hostemu


def test_rings_are_vertex_stars():
    m = sp.MeshTri()
    m.refine(3)
    nb, nfan, closed = vertex_rings(m)
    nv = m.p.shape[1]
    assert np.array_equal(nfan, np.bincount(m.t.reshape(-1), minlength=nv))
    bnd = np.zeros(nv, dtype=bool)
    bnd[m.boundary_nodes()] = True
    assert np.array_equal(closed, ~bnd)
    # neighbours of v = vertices sharing an edge with v
    for v in (0, 5, 17, nv - 1):
        ring = set(nb[:, v][nb[:, v] >= 0].tolist())
        edges = m.facets[:, (m.facets == v).any(axis=0)]
        assert ring == set(edges.reshape(-1).tolist()) - {v}
    # counter-clockwise order: consecutive neighbours span positively oriented triangles
    v = int(np.nonzero(closed)[0][0])
    r = nb[:nfan[v], v]
    for k in range(nfan[v]):
        a, b = m.p[:, r[k]] - m.p[:, v], m.p[:, r[(k + 1) % nfan[v]]] - m.p[:, v]
        assert a[0] * b[1] - a[1] * b[0] > 0


@pytest.mark.parametrize("tiny", [True, False])
@pytest.mark.parametrize("name", ["trip1_lap", "trip1_mass", "trip1_lap_r6"])
def test_fan_kernel_matches_reference(name, tiny):
    case = cases.BY_NAME[name]
    asm = util.assembler_for(case)
    out, fp = hostemu.fan_iasm(asm, case.form, rows_per_tile=64, tiny=tiny)
    assert fp["format"] == ("tiny" if tiny else "fan8")   # 20- / 32-byte row records
    assert not np.isnan(out.data).any()            # every CSR value written
    util.check(out, util.golden(case))
    assert fp["halo_factor"] < 2.0


def test_fan_kernel_on_jiggled_symmetric_mesh():
    """Irregular valences (4..8) and a form using h."""
    m = sp.MeshTri(initmesh="symmetric")
    m.refine(3)
    rng = np.random.default_rng(1)
    I = m.interior_nodes()
    m.p[:, I] += 0.01 * rng.standard_normal((2, len(I)))
    asm = AssemblerElement(m, sp.ElementTriP1())
    form = lambda u, v, du, dv, h: h * (du[0] * dv[0] + 2.0 * du[1] * dv[1]) + u * v + du[0] * v
    out, _ = hostemu.fan_iasm(asm, form, rows_per_tile=32)
    ref = hostemu.iasm(asm, form)
    cases.compare_csr(out, ref)


def test_fan_plan_rejects_other_patterns():
    m = sp.MeshTri()
    m.refine(2)
    asm = AssemblerElement(m, sp.ElementTriP2())
    A = hostemu.iasm(asm, cases.mass)
    with pytest.raises(FanPlanError):
        build_fan_plan(m, A.indptr, A.indices)
    assert not asm.plan_iasm(cases.mass).spec.fan_ok
    a1 = AssemblerElement(m, sp.ElementTriP1())
    assert not a1.plan_iasm(cases.mass_x2).spec.fan_ok      # x-dependent coefficient


def _wheel_mesh(k=10, refine=2):
    """A vertex with ``k`` triangles around it (the wide, 12-neighbour record
    layout and kernel instantiation)."""
    ang = 2 * np.pi * np.arange(k) / k
    p = np.hstack((np.zeros((2, 1)), np.vstack((np.cos(ang), np.sin(ang)))))
    t = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)]).T
    m = sp.MeshTri(p, t)
    m.refine(refine)
    return m


def test_fan_kernel_wide_records():
    m = _wheel_mesh()
    asm = AssemblerElement(m, sp.ElementTriP1())
    out, fp = hostemu.fan_iasm(asm, cases.lap2, rows_per_tile=16)
    assert fp["max_neighbours"] == 10
    cases.compare_csr(out, hostemu.iasm(asm, cases.lap2))


@pytest.mark.parametrize("tiny", [True, False])
@pytest.mark.parametrize("which", ["refined", "wheel", "jiggled", "clockwise"])
def test_native_plan_is_byte_identical(which, tiny):
    """The native plan builder (``spfem_fanplan_build``, csrc/spfem_plan.cu; run
    here on host pointers, the same code as on the device) produces exactly the
    blob / descriptors of the NumPy reference builder."""
    import torch
    from spfem_b200 import fanplan
    if which == "refined":
        m = sp.MeshTri()
        m.refine(4)
    elif which == "wheel":
        m = _wheel_mesh(10, 2)
    elif which == "clockwise":          # some triangles listed clockwise
        m = sp.MeshTri()
        m.refine(3)
        t = m.t.copy()
        t[[1, 2], ::3] = t[[2, 1], ::3]
        m = sp.MeshTri(m.p, t)
    else:
        m = sp.MeshTri(initmesh="symmetric")
        m.refine(3)
        I = m.interior_nodes()
        m.p[:, I] += 0.01 * np.random.default_rng(1).standard_normal((2, len(I)))
    asm = AssemblerElement(m, sp.ElementTriP1())
    A = hostemu.iasm(asm, cases.mass)
    a = build_fan_plan(m, A.indptr, A.indices, 32, tiny=tiny)
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    for idt in (np.int32, np.int64):
        b = fanplan.build_fan_plan(torch, T(m.p), T(m.t.astype(np.int32)), T(A.indptr.astype(idt)),
                                   T(A.indices.astype(idt)), 32, tiny=tiny)
        assert a["format"] == b["format"] and (a["format"] == "tiny") == (tiny and which != "wheel")
        assert np.array_equal(a["desc"], b["desc"].numpy())
        assert np.array_equal(a["vid"], b["vid"].numpy())
        assert np.array_equal(a["pc_index"], b["pc_index"].numpy())
        assert np.array_equal(a["blob"], b["blob"].numpy())
        for k in ("ntiles", "max_a", "max_ns", "max_neighbours"):
            assert a[k] == b[k]
        assert abs(a["halo_factor"] - b["halo_factor"]) < 1e-12


def test_native_plan_rejects_what_the_reference_builder_rejects():
    import torch
    from spfem_b200 import fanplan
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    # a vertex with 13 triangles around it
    m = _wheel_mesh(13, 0)
    asm = AssemblerElement(m, sp.ElementTriP1())
    A = hostemu.iasm(asm, cases.mass)
    with pytest.raises(FanPlanError):
        build_fan_plan(m, A.indptr, A.indices)
    with pytest.raises(fanplan.FanPlanError):
        fanplan.build_fan_plan(torch, T(m.p), T(m.t.astype(np.int32)), T(A.indptr.astype(np.int32)),
                               T(A.indices.astype(np.int32)))
    # a pattern that is not the vertex-star pattern (P2 matrix on the P1 mesh)
    m = sp.MeshTri()
    m.refine(2)
    A2 = hostemu.iasm(AssemblerElement(m, sp.ElementTriP2()), cases.mass)
    with pytest.raises(fanplan.FanPlanError):
        fanplan.build_fan_plan(torch, T(m.p), T(m.t.astype(np.int32)),
                               T(A2.indptr[:m.p.shape[1] + 1].astype(np.int32)),
                               T(A2.indices.astype(np.int32)))
