# -*- coding: utf-8 -*-
"""Chunked single-GPU assembly (meshes beyond one pattern call) on the GPU.  Its
on-device merge was rewritten after round 1's GPU minutes were spent (CPU
tensors only so far, tests/test_host_logic.py), so the test sits in a file that
runs last."""
import pytest

import cases
import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["trip1_lap_r6", "tetp2_lap", "vecp2_elasticity", "stokes_b",
                                  "trip1_load"])
def test_chunked_assembly_beyond_one_pattern_call(name, monkeypatch):
    """``iasm`` on a mesh with more local entries than one pattern call takes
    (limit lowered through SPFEM_MAX_ENTRIES): element chunks on one GPU, same
    matrix and pattern as the single call."""
    case = cases.BY_NAME[name]
    asm = util.assembler_for(case)
    plan = asm.plan_iasm(case.form)
    monkeypatch.setenv("SPFEM_MAX_ENTRIES", str(plan.spec.n_entries * asm.mesh.t.shape[1] // 3))
    assert asm._chunks_needed(plan, None) >= 4
    util.check(asm.iasm(case.form), util.golden(case))
