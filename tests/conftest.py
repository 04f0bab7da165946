import os
import sys
import warnings

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
warnings.filterwarnings("ignore", category=SyntaxWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _gpu_ready():
    """A CUDA device and the built C-ABI library (the product has no CPU path)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as exc:                      # pragma: no cover
        return False, "torch: %s" % exc
    lib = os.path.join(ROOT, "spfem_b200", "csrc", "libspfem_b200.so")
    if not os.path.exists(lib):
        return False, "libspfem_b200.so is not built"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """``gpu``-marked tests are skipped (not failed) on a box without a GPU, so a
    plain ``pytest tests`` separates host-side regressions from the expected
    'no device' condition."""
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason="gpu test: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
