# -*- coding: utf-8 -*-
"""Tracer + code generator checked on the CPU: the generated element / facet
kernels are compiled for the host (``-DSPFEM_HOST``, see tests/hostemu.py) and
compared with the reference's golden outputs.  No GPU, no CUDA calls."""
import numpy as np
import pytest

import cases
import util
import hostemu


@pytest.mark.parametrize("case", cases.CASES, ids=lambda c: c.name)
def test_generated_kernel_matches_reference(case):
    asm = util.assembler_for(case)
    kw = cases.case_inputs(case, asm.mesh, asm.dofnum_u.N)
    fn = hostemu.iasm if case.kind == "iasm" else hostemu.fasm
    util.check(fn(asm, case.form, **kw), util.golden(case))
