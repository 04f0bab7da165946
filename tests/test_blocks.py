# -*- coding: utf-8 -*-
"""Block systems on the device (spfem_b200/blocks.py) against SciPy: stacking
(spfem/utils.py:12-20), Dirichlet elimination and solve (:52-95).  The CPU
suite runs the tensor code on CPU tensors with matrices from the host
emulation; ``-m gpu`` runs the whole chain on the device."""
import numpy as np
import pytest
import scipy.sparse as spsp
import torch

import spfem_b200 as sp
from spfem_b200 import blocks
from spfem_b200.assembly import AssemblerElement
import cases
import hostemu


def _stokes_blocks(assemble):
    m = sp.MeshTri()
    m.refine(3)
    a = AssemblerElement(m, sp.ElementH1Vec(sp.ElementTriP2()))
    b = AssemblerElement(m, sp.ElementH1Vec(sp.ElementTriP2()), sp.ElementTriP1())
    c = AssemblerElement(m, sp.ElementTriP1())
    return (assemble(a, cases.elasticity), assemble(b, cases.stokes_b),
            assemble(c, cases.mass))


def test_stack_and_transpose_match_scipy():
    A, B, C = _stokes_blocks(hostemu.iasm)
    K = blocks.stack(torch, [[A, ("T", B)], [B, None]])
    ref = spsp.vstack([spsp.hstack([A, B.T]), spsp.hstack([B, 0 * C])]).tocsr()
    ref.sort_indices()
    got = K.to_scipy()
    assert got.shape == ref.shape
    assert abs(got - ref).max() == 0.0
    # the explicit zeros of the assembled blocks are kept, the zero block adds none
    assert got.nnz == A.nnz + 2 * B.nnz
    Bt = blocks.transpose(torch, B).to_scipy()
    assert abs(Bt - B.T.tocsr()).max() == 0.0 and Bt.has_sorted_indices
    K2 = blocks.stack(torch, [[blocks.as_csr(torch, A), blocks.as_csr(torch, B).T],
                              [B, None]])
    assert abs(K2.to_scipy() - ref).max() == 0.0
    with pytest.raises(ValueError):
        blocks.stack(torch, [[A, None], [None, None]])


def test_condense_and_cg_match_direct():
    m = sp.MeshTri()
    m.refine(4)
    a = AssemblerElement(m, sp.ElementTriP1())
    K = hostemu.iasm(a, cases.lap2)
    f = hostemu.iasm(a, cases.load1)
    I = m.interior_nodes()
    D = m.boundary_nodes()
    g = np.zeros(K.shape[0])
    g[D] = 1.0 + m.p[0, D]                        # inhomogeneous Dirichlet data
    ref = sp.direct(K, f, x=g.copy(), I=I)
    Kd, fd = blocks.as_csr(torch, K), torch.from_numpy(f)
    AII, bI, It = blocks.condense(torch, Kd, fd, I, torch.from_numpy(g))
    assert abs(AII.to_scipy() - K[I].T[I].T).max() == 0.0
    assert np.allclose(bI.numpy(), f[I] - K[I].T[D].T.dot(g[D]), rtol=0, atol=1e-14)
    x = blocks.solve_dirichlet(torch, Kd, fd, I, torch.from_numpy(g), tol=1e-12)
    assert np.max(np.abs(x.numpy() - ref)) < 1e-9
    xh, it, res = blocks.cg(torch, AII, bI, tol=1e-12)
    assert res <= 1e-12 and it < AII.shape[0]


@pytest.mark.gpu
def test_poisson_solved_on_the_device():
    """README example with the matrix, the elimination and the solve on the GPU
    (known answer of spfem/test_asm.py:163)."""
    m = sp.MeshTri()
    m.refine(5)
    a = AssemblerElement(m, sp.ElementTriP1())
    K = a.iasm_device(cases.lap2)
    f = a.iasm_device(cases.load1)
    x = blocks.solve_dirichlet(torch, K, f, m.interior_nodes(), tol=1e-12)
    assert x.is_cuda
    assert abs(float(x.max()) - 0.073614737354524146) < 1e-9


@pytest.mark.gpu
def test_stokes_blocks_stacked_on_the_device():
    dev = lambda asm, form: asm.iasm_device(form)
    A, B, C = _stokes_blocks(dev)
    K = blocks.stack(torch, [[A, ("T", B)], [B, None]])
    assert K.values.is_cuda
    Ah, Bh = A.to_scipy(), B.to_scipy()
    ref = spsp.vstack([spsp.hstack([Ah, Bh.T]), spsp.hstack([Bh, 0 * C.to_scipy()])]).tocsr()
    assert abs(K.to_scipy() - ref).max() == 0.0
