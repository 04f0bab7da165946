#!/usr/bin/env python
"""Host vs device topology builder on the benchmark meshes (equal tables, time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import spfem_b200 as sp
torch.zeros(1, device="cuda")
for cls, ref in (("MeshTri", 11), ("MeshTet", 6)):
    res = {}
    for mode in ("0", "auto"):
        os.environ["SPFEM_DEVICE_TOPOLOGY"] = mode
        t0 = time.perf_counter()
        m = getattr(sp, cls)(); m.refine(ref)
        res[mode] = (time.perf_counter() - t0, m)
    a, b = res["0"][1], res["auto"][1]
    same = all(np.array_equal(getattr(a, n), getattr(b, n))
               for n in ("p", "t", "facets", "t2f", "f2t") + (("edges", "t2e") if a._local_edges else ()))
    print(cls, ref, "nt", a.t.shape[1], "host %.2f s  device %.2f s  identical %s"
          % (res["0"][0], res["auto"][0], same))
