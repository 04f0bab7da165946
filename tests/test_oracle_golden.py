# -*- coding: utf-8 -*-
"""Pin the oracle: it must reproduce the reference's own outputs (golden
fixtures generated from /root/reference by tests/golden/make_golden.py) and the
known answers the reference's tests hold for the path."""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

import cases
import util
import asm_oracle


@pytest.mark.parametrize("case", [c for c in cases.CASES if c.oracle], ids=lambda c: c.name)
def test_oracle_matches_reference_output(case):
    util.check(util.oracle_run(case), util.golden(case))


def test_known_answer_test_asm_163():
    """spfem/test_asm.py:145-163: P1 Poisson on refine(5), max(x) to 7 places."""
    import spfem_b200 as sp
    m = sp.MeshTri()
    m.refine(5)
    A = asm_oracle.iasm(m, "TriP1", cases.lap2)
    b = asm_oracle.iasm(m, "TriP1", cases.load1)
    I = m.interior_nodes()
    x = np.zeros(A.shape[0])
    x[I] = spl.spsolve(A[I].T[I].T.tocsc(), b[I])
    assert abs(np.max(x) - 0.073614737354524146) < 1e-7


def test_survey_fingerprints():
    """SURVEY.md 8(c) fingerprints of the patched reference (TriP1 r6)."""
    import hashlib
    A = util.golden(cases.BY_NAME["trip1_lap_r6"])
    pat = hashlib.sha256(A.indptr.astype(np.int64).tobytes()
                         + A.indices.astype(np.int64).tobytes()).hexdigest()[:16]
    assert A.shape == (4225, 4225) and A.nnz == 29057
    assert pat == "3411fe8c348d15b1"
    assert int(np.sum(A.data == 0.0)) == 8192
    B = util.oracle_run(cases.BY_NAME["trip1_lap_r6"])
    assert int(np.sum(B.data == 0.0)) >= 8000   # explicit zeros are kept
