# -*- coding: utf-8 -*-
"""Vertex-fan plan (``codegen.FAN_KERNEL``): Python side of the native builder
``spfem_fanplan_build`` (csrc/spfem_plan.cu, declared in include/spfem_b200.h).

Scalar P1 on triangles with element-constant coefficients: one thread per CSR
row = mesh vertex ``r`` walks the triangles around ``r`` in counter-clockwise
ring order.  The plan holds, per tile of ``rows_per_tile`` Morton-consecutive
vertices (rows of a tile in CSR order), one 16-byte aligned blob:

  coordinates  f64 [nvt][2]   interleaved (x, y) of the tile's vertices: its rows
                              first, then the ring neighbours owned by other tiles
                              (halo, in id order); a private copy, refreshed by
                              :func:`fill_fan_coordinates`
  row records  SoA planes of 16 bytes per row:
      plane 0  {CSR position of the row, s0 | nfan << 16 | closed << 20 | diagonal
                position << 24, neighbours 0..3}
      format "tiny" (<= 8 ring neighbours, <= 256 vertices per tile; 20 bytes per
                row): neighbours as 8-bit tile-local ids in plane 0, one extra
                32-bit word per row with the 4-bit positions of the neighbours
                inside the CSR row
      format "fan8" (<= 8 neighbours; 32 bytes): 16-bit ids, plane 1 {neighbours
                4..7, 8-bit positions 0..7}
      format "wide" (<= 12 neighbours; 48 bytes): plane 1 {neighbours 4..11},
                plane 2 {positions 0..11}

Descriptor (8 int32 per tile): blob offset / 16, bytes of the coordinates, bytes
of the records, rows, CSR slots, vertices, 0, 0.  ``tests/fanplan_ref.py`` is the
NumPy statement of the same plan (byte-identical; ``tests/test_fanplan.py``)."""
import ctypes as C

import numpy as np

from . import _lib
from .tileplan import _Arena

FORMATS = ("tiny", "fan8", "wide")


class FanPlanError(ValueError):
    """The mesh / pattern is not vertex-fan shaped (the caller falls back to the
    element-tile kernel)."""


def build_fan_plan(torch, p, t, indptr, indices, rows_per_tile=128, tiny=True,
                   row_order="id", stream=None):
    """``p`` (2, nv) float64, ``t`` (3, nt) int32, ``indptr`` / ``indices`` of the
    vertex x vertex CSR pattern (int32 or int64), all on one device (CPU tensors
    run the same code on the host: tests).  Returns the plan as a dict of tensors;
    raises :class:`FanPlanError` when the vertex-fan kernel does not apply."""
    dev = p.device
    on_dev = dev.type == "cuda"
    lib = _lib.load()
    if p.shape[0] != 2 or indptr.numel() != p.shape[1] + 1:
        raise FanPlanError("not a scalar vertex-based pattern on a 2-D mesh")
    p = p.contiguous()
    t = t.to(torch.int32).contiguous()
    indptr, indices = indptr.contiguous(), indices.contiguous()
    if indptr.dtype != indices.dtype or indptr.dtype not in (torch.int32, torch.int64):
        raise FanPlanError("index arrays must both be int32 or int64")
    pin = _lib.SpfemFanPlanIn()
    pin.p, pin.ldp, pin.nv = p.data_ptr(), p.shape[1], p.shape[1]
    pin.t, pin.ldt, pin.nt = t.data_ptr(), t.shape[1], t.shape[1]
    pin.indptr, pin.indices = indptr.data_ptr(), indices.data_ptr()
    pin.index_bytes = 8 if indptr.dtype == torch.int64 else 4
    pin.rows_per_tile, pin.tiny = int(rows_per_tile), 1 if tiny else 0
    pin.row_order = 0 if row_order == "id" else 1
    lo, hi = p.min(dim=1).values.tolist(), p.max(dim=1).values.tolist()
    for k in range(2):
        pin.lo[k], pin.hi[k] = lo[k], hi[k]
    pout = _lib.SpfemFanPlanOut()
    arena = _Arena(torch, dev)
    if stream is None and on_dev:
        stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.spfem_fanplan_build(C.byref(pin), C.byref(pout), arena.alloc_cb, arena.free_cb,
                                 None, 1 if on_dev else 0, C.c_void_p(stream or 0))
    if rc != 0:
        msg = lib.spfem_last_error().decode("utf-8", "replace")
        arena.live.clear()
        if rc == 5:
            raise FanPlanError(msg)
        if rc == 3 and arena.error is not None:
            raise arena.error
        raise _lib.SpfemError(msg)
    nt_ = int(pout.ntiles)
    plan = {
        "ntiles": nt_, "rows_per_tile": int(rows_per_tile),
        "desc": arena.tensor(pout.desc, torch.int32, 8 * nt_).view(nt_, 8),
        "blob": arena.tensor(pout.blob, torch.uint8, int(pout.blob_bytes)),
        "vid": arena.tensor(pout.vid, torch.int32, int(pout.nvid)),
        "pc_index": arena.tensor(pout.pc_index, torch.int64, int(pout.nvid)),
        "max_a": int(pout.max_a), "max_b": 0, "max_ns": int(pout.max_ns),
        "halo_factor": float(pout.halo_factor),
        "max_neighbours": int(pout.max_neighbours),
        "format": FORMATS[int(pout.format)],
        "scratch_peak_bytes": arena.peak,
    }
    arena.live.clear()
    return plan


def fill_fan_coordinates(plan, p):
    """Copy the vertex coordinates ``p`` (2, nv) into the tiles' private
    interleaved (x, y) sections (call after every change of ``p``)."""
    b64 = plan["blob"].view(p.dtype)
    vid = plan["vid"].long()
    for d in range(2):
        b64[plan["pc_index"] + d] = p[d][vid]
