# -*- coding: utf-8 -*-
"""``import spfem`` on top of spfem_b200: put ``<repo>/compat`` on ``PYTHONPATH``
and scripts written for the reference (``from spfem.mesh import MeshTri``,
``from spfem.asm import AssemblerElement``, ``import spfem; spfem.MeshTri()``)
run on the B200 path unchanged.  Mirrors spfem/__init__.py:1-6 and the
backwards-compatibility alias spfem/asm.py:7.  (Kept out of the repository root
on purpose: the test oracle imports the patched reference under the same
top-level name.)"""
import importlib
import sys

import spfem_b200 as _impl
from spfem_b200 import *                                   # noqa: F401,F403
from spfem_b200.assembly import Assembler, AssemblerElement, Dofnum  # noqa: F401

for _name in ("mesh", "assembly", "element", "mapping", "utils", "quadrature"):
    _mod = importlib.import_module("spfem_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
sys.modules[__name__ + ".asm"] = sys.modules[__name__ + ".assembly"]
asm = sys.modules[__name__ + ".asm"]

__all__ = ['MeshTri', 'MeshTet', 'MeshQuad', 'MeshLine', 'AssemblerElement',
           'ElementTriP1', 'ElementTetP1', 'ElementTriDG', 'ElementTetDG', 'direct']
__version__ = _impl.__version__
