# -*- coding: utf-8 -*-
"""spfem_b200 -- B200-native element assembly behind the sp.fem API.

Import surface mirrors ``spfem/__init__.py``: meshes, elements, the
``AssemblerElement`` (device assembly) and the host helpers."""
from .mesh import Mesh, MeshTri, MeshTet, MeshQuad, MeshLine
from .element import (Element, ElementH1, ElementH1Vec, ElementTriP1,
                      ElementTriP2, ElementTetP1, ElementTetP2, ElementQ1,
                      ElementQ2, ElementTriP0, ElementP0, ElementTetP0,
                      ElementTriMini, ElementTriDG, ElementTetDG, ElementTriPp,
                      ElementHdiv, ElementTriRT0)
from .utils import direct, stack, const_cell, cell_shape
from .assembly import Assembler, AssemblerElement, Dofnum

__version__ = "0.1.0"
