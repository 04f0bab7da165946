# -*- coding: utf-8 -*-
"""Mesh topology tables: the torch (device) builders of spfem_b200/mesh.py run on
CPU tensors must reproduce the NumPy builders bit for bit (facets / edges in
lexicographic order, t2f / t2e, f2t with the first / last occurrence rule); on
a GPU the same check runs through ``Mesh._build_mappings_device``."""
import numpy as np
import pytest
import torch

import spfem_b200 as sp
from spfem_b200 import mesh as M


def _meshes():
    out = []
    m = sp.MeshTri(); m.refine(3); out.append(m)
    m = sp.MeshTet(); m.refine(2); out.append(m)
    m = sp.MeshQuad(); m.refine(2); out.append(m)
    m = sp.MeshTri(); m.refine(2); m.p = m.p + 0.01 * np.sin(37.0 * m.p[::-1]); out.append(m)
    return out


@pytest.mark.parametrize("mesh", _meshes(), ids=lambda m: type(m).__name__)
def test_torch_topology_matches_numpy(mesh):
    t = torch.from_numpy(np.ascontiguousarray(mesh.t, dtype=np.int64))
    nv = mesh.p.shape[1]
    f, t2f = M._entity_table_torch(torch, t, mesh._local_facets, nv)
    assert np.array_equal(f.numpy(), mesh.facets) and np.array_equal(t2f.numpy(), mesh.t2f)
    f2t = M._facet_to_element_torch(torch, t2f, f.shape[1])
    assert np.array_equal(f2t.numpy(), mesh.f2t)
    if mesh._local_edges:
        e, t2e = M._entity_table_torch(torch, t, mesh._local_edges, nv)
        assert np.array_equal(e.numpy(), mesh.edges) and np.array_equal(t2e.numpy(), mesh.t2e)


def test_device_topology_switch(monkeypatch):
    monkeypatch.setenv("SPFEM_DEVICE_TOPOLOGY", "0")
    assert not M._device_topology_wanted(1 << 20)
    monkeypatch.setenv("SPFEM_DEVICE_TOPOLOGY", "auto")
    assert not M._device_topology_wanted(1000)          # small meshes stay on the host


@pytest.mark.gpu
@pytest.mark.parametrize("cls,ref", [("MeshTri", 6), ("MeshTet", 3), ("MeshQuad", 5)])
def test_device_topology_on_gpu(cls, ref, monkeypatch):
    monkeypatch.setenv("SPFEM_DEVICE_TOPOLOGY", "0")
    a = getattr(sp, cls)(); a.refine(ref)
    monkeypatch.setenv("SPFEM_DEVICE_TOPOLOGY", "1")
    b = getattr(sp, cls)(); b.refine(ref)
    for name in ("p", "t", "facets", "t2f", "f2t") + (("edges", "t2e") if a._local_edges else ()):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert getattr(a, name).dtype == getattr(b, name).dtype, name
