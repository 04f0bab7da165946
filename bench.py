#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Headline benchmark: elements assembled per second into a global fp64 CSR.

Workload (BASELINE.json configs[1]): ``MeshTri().refine(11)`` (8 388 608
triangles, 4 198 401 vertices), ``ElementTriP1``, one *step* = assemble the
Laplace stiffness matrix AND the mass matrix (2 x nt element assemblies) from
the device-resident mesh into device-resident CSR values.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--refine R]
    python bench.py --impl reference ...      # reference CPU arm

Prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "elements assembled/s to global CSR (fp64)"
UNIT = "elements/s"


def lap2(du, dv):
    return du[0] * dv[0] + du[1] * dv[1]


def mass(u, v):
    return u * v


FORMS = (("laplace", lap2), ("mass", mass))


def algorithmic_bytes(nt, nv, nnz, nve=3, dim=2):
    """connectivity ints + each vertex once + each CSR value
    once (TriP1: 12*nt + 16*nv + 8*nnz)."""
    return 4 * nve * nt + 8 * dim * nv + 8 * nnz


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index=0, period=0.002):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.recording = False        # samples count only inside the timed region
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.recording:
                    self.samples.append(mhz)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self.recording = False
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# reference arm / CPU baseline
# ---------------------------------------------------------------------------
def _reference_modules():
    """The patched reference (oracle/_ref) if it travelled with the repo,
    else None (then the numpy oracle port is timed)."""
    try:
        import make_ref
        ref = make_ref.import_ref()
        if ref is None:
            return None
        import spfem.mesh as rmesh
        import spfem.element as relem
        import spfem.assembly as rasm
        return rmesh, relem, rasm
    except Exception:
        return None


def cpu_assemble_factory(refine):
    """Returns ``(step, nt, kind)``: ``step()`` assembles Laplace + mass with
    the reference's own CPU implementation on a refine(``refine``) mesh."""
    mods = _reference_modules()
    if mods is not None:
        rmesh, relem, rasm = mods
        m = rmesh.MeshTri()
        m.refine(refine)
        a = rasm.AssemblerElement(m, relem.ElementTriP1())

        def step(best_case=False):
            # best_case: the same API call with tind=np.arange(nt) (fancy instead of
            # Python-range indexing inside the reference; SURVEY.md 8(d), identical matrix)
            tind = np.arange(m.t.shape[1]) if best_case else None
            for _, f in FORMS:
                a.iasm(f, tind=tind)
        return step, m.t.shape[1], "reference"
    import spfem_b200 as sp
    import asm_oracle
    m = sp.MeshTri()
    m.refine(refine)

    def step(best_case=False):
        for _, f in FORMS:
            asm_oracle.iasm(m, "TriP1", f)
    return step, m.t.shape[1], "port"


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the reference arm runs a larger sample than the baseline beside the GPU line
    # (refine(10) = 2 M triangles: the reference's throughput still depends on the
    # mesh size below that), bounded in time: at least one timed step, then stop
    # after --ref-seconds
    refine = args.ref_refine
    step, nt, kind = cpu_assemble_factory(refine)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        step()
        done += 1
        if time.perf_counter() - t0 > args.ref_seconds:
            break
    dt = time.perf_counter() - t0
    value = 2.0 * nt * done / dt
    args.steps = done
    sample = ("MeshTri().refine(%d): %d triangles x 2 forms per step, %d timed steps, %s "
              "AssemblerElement.iasm as shipped (tind=None)" % (refine, nt, done,
              "patched reference" if kind == "reference" else "numpy oracle port"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "MeshTri.refine(11) ElementTriP1 Laplace+mass iasm "
                               "(CPU arm: bounded sample, see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(),
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# workloads of BASELINE.json configs 3 / 4 / 5 and the target-size TriP1 mesh
# ---------------------------------------------------------------------------
def _eps(dw):
    return [[dw[0][0], 0.5 * (dw[0][1] + dw[1][0])],
            [0.5 * (dw[0][1] + dw[1][0]), dw[1][1]]]


def elasticity(du, dv):
    # 2 mu eps(u):eps(v) + lambda tr eps(u) tr eps(v), lambda = mu = 1
    # (tmp/elastic_contact.py:49-72 in the reference, written out)
    eu, ev = _eps(du), _eps(dv)
    return (2.0 * sum(eu[i][j] * ev[i][j] for i in range(2) for j in range(2))
            + 1.0 * (eu[0][0] + eu[1][1]) * (ev[0][0] + ev[1][1]))


def stokes_a(du, dv):
    # sym-grad form of the Stokes notebook (learning/Example 1, cell 9)
    eu, ev = _eps(du), _eps(dv)
    return sum(eu[i][j] * ev[i][j] for i in range(2) for j in range(2))


def stokes_b(du, v):
    return (du[0][0] + du[1][1]) * v


def lap3(du, dv):
    return du[0] * dv[0] + du[1] * dv[1] + du[2] * dv[2]


def vec_nitsche(u, v, du, dv, h, n):
    # gamma/h u.v - (grad u n).v - u.(grad v n)   (vector version of
    # spfem/test_module.py:500-501)
    gamma = 100.0
    dun = [du[i][0] * n[0] + du[i][1] * n[1] for i in range(2)]
    dvn = [dv[i][0] * n[0] + dv[i][1] * n[1] for i in range(2)]
    return sum(gamma / h * u[i] * v[i] - dun[i] * v[i] - u[i] * dvn[i] for i in range(2))


# name: (grid kind, default size, C = connectivity ints per element, description)
CONFIGS = {
    "trip1_refine13": ("refine", 13, 3, "MeshTri.refine(13) ElementTriP1 Laplace (target size: >= 50 M elements per GPU)"),
    "trip1_grid": ("tri", 5792, 3, "n x n grid (67 M triangles, row-major vertex numbering), ElementTriP1 Laplace"),
    "vecp2_elasticity": ("tri", 3162, 6, "config 3: n x n grid, ElementH1Vec(ElementTriP2) linear elasticity"),
    "stokes_A": ("tri", 3162, 6, "config 5: Taylor-Hood velocity block, vector-P2 sym-grad form"),
    "stokes_B": ("tri", 3162, 6, "config 5: divergence block, rows P1 x columns vector-P2"),
    "stokes_C": ("tri", 3162, 3, "config 5: pressure mass block, ElementTriP1"),
    "nitsche_fasm": ("tri", 3162, 0, "config 5: boundary fasm, Nitsche terms on vector-P2 (4 n facets)"),
    "tetp2_laplace": ("tet", 203, 10, "config 4: Kuhn m^3 grid, ElementTetP2 Laplace"),
    "q2_laplace": ("quadrefine", 11, 4, "MeshQuad.refine(11) ElementQ2 Laplace (MappingQ1: quadrature-point loop)"),
}


_MESH_CACHE = {}        # the config-5 blocks share one grid


def _time_launches(torch, preps, steps, barrier):
    """ms per launch of every prepared assembly (CUDA events on the launching
    stream, after 3 warm-up launches) and the number of kernel launches."""
    for _ in range(3):
        for p in preps:
            p.launch()
    barrier()
    out = []
    for p in preps:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            p.launch()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
    return out


def run_config(name, args, world, rank, peak, traffic, barrier):
    """One entry of the ``configs`` array: builds the mesh (N > 1: this rank's
    slab of the weak-scaled grid, spfem_b200/distributed.py::slab_part), stages
    the form(s), times the kernels.  Returns a dict (rank 0 reports it)."""
    import torch
    import torch.distributed as dist
    import spfem_b200 as sp
    from spfem_b200 import meshgen
    from spfem_b200.assembly import AssemblerElement
    from spfem_b200.distributed import DistributedAssembler, slab_part
    kind, size, C, desc = CONFIGS[name]
    size = int(args.sizes.get(name, size))
    out = {"name": name, "workload": desc, "size": size}
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    vecp2 = lambda: sp.ElementH1Vec(sp.ElementTriP2())
    spec = {
        "trip1_refine13": (sp.ElementTriP1, None, lap2, "iasm"),
        "trip1_grid": (sp.ElementTriP1, None, lap2, "iasm"),
        "vecp2_elasticity": (vecp2, None, elasticity, "iasm"),
        "stokes_A": (vecp2, None, stokes_a, "iasm"),
        "stokes_B": (vecp2, sp.ElementTriP1, stokes_b, "iasm"),
        "stokes_C": (sp.ElementTriP1, None, mass, "iasm"),
        "nitsche_fasm": (vecp2, None, vec_nitsche, "fasm"),
        "tetp2_laplace": (sp.ElementTetP2, None, lap3, "iasm"),
        "q2_laplace": (sp.ElementQ2, None, lap2, "iasm"),
    }[name]
    eu = spec[0]()
    ev = spec[1]() if spec[1] is not None else None
    form, how = spec[2], spec[3]
    part = None
    if kind in ("refine", "quadrefine"):
        if world > 1:
            return dict(out, skipped="single-GPU configuration")
        m = sp.MeshTri() if kind == "refine" else sp.MeshQuad()
        m.refine(size)
        nt_own = m.t.shape[1]
    elif world == 1:
        m = _MESH_CACHE.get((kind, size))
        if m is None:
            _MESH_CACHE.clear()
            p, t = (meshgen.grid_tri(size) if kind == "tri" else meshgen.grid_tet(size))
            m = (sp.MeshTri if kind == "tri" else sp.MeshTet)(p, t, validate=False)
            _MESH_CACHE[(kind, size)] = m
        nt_own = m.t.shape[1]
    else:
        part = slab_part(kind, size, rank, world, eu, ev)
        m = part.mesh
        nt_own = part.n_own_elements
    out["mesh_s"] = round(time.perf_counter() - t0, 2)
    t0 = time.perf_counter()
    if part is None:
        a = AssemblerElement(m, eu, ev)
    else:
        da = DistributedAssembler(None, eu, ev, part=part, rank=rank, world=world)
        a = da.local
    out["dofnum_s"] = round(time.perf_counter() - t0, 2)
    t0 = time.perf_counter()
    if how == "iasm":
        prep = a.prepare_iasm(form)
        n_units = nt_own
    else:
        find = m.boundary_facets() if part is None else part.global_boundary_facets()
        prep = a.prepare_fasm(form, find=find)
        n_units = len(find)
    torch.cuda.synchronize()
    out["prepare_s"] = round(time.perf_counter() - t0, 2)
    out["first_call_s"] = round(out["dofnum_s"] + out["prepare_s"], 2)
    steps = max(3, min(args.steps, args.config_steps))
    ms = _time_launches(torch, [prep], steps, barrier)[0]
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    nv, dim = m.p.shape[1], m.p.shape[0]
    pat = prep.pattern
    nnz = pat.nnz if prep.plan.bilinear else pat.nrows
    out.update({
        "route": type(prep).__name__, "launches_per_assembly": prep.n_launches,
        "units": "elements" if how == "iasm" else "boundary facets",
        "n_per_gpu": int(n_units), "n_total": int(n_units) * world, "nv": int(nv),
        "nnz_per_gpu": int(nnz), "ms_per_assembly": ms,
        "value": n_units * world / (ms * 1e-3), "unit": UNIT if how == "iasm" else "facets/s",
        "steps": steps,
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
    })
    if how == "iasm":
        balg = algorithmic_bytes(nt_own, nv, nnz, nve=C, dim=dim)
        ach = balg / (ms * 1e-3) / 1e9
        tile = None
        if hasattr(prep, "tp"):
            tp = prep.tp
            tile = {k: tp.get(k) for k in ("kind", "E", "ntiles", "halo_factor", "bsv", "bsu",
                                           "compact", "sort_slots", "blob_bytes", "tiles_per_cta")
                    if tp.get(k) is not None}
            tile["smem_bytes"] = prep.smem
            tile["threads"] = prep.threads
        # DRAM bytes per launch: ncu capture of the same kernel on a smaller grid of the
        # same family, scaled by the number of elements (profiles/traffic_r02.json)
        tr = traffic.get(name)
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak,
                           "traffic": (int(tr["dram_bytes_per_element"] * nt_own) if tr else None),
                           "traffic_source": (("ncu on %s, %.1f B/element x elements"
                                               % (tr["measured_on"], tr["dram_bytes_per_element"]))
                                              if tr else None),
                           "algorithmic_bytes_per_assembly": balg,
                           "algorithmic_bytes_per_element": round(balg / nt_own, 1),
                           "tile": tile}
    else:
        out["note"] = ("O(sqrt(nt)) work through the two-kernel route (facet kernel + "
                       "fixed-order reduction): launch-latency bound, no roofline claimed")
    del prep, a, m
    torch.cuda.empty_cache()
    return out


def multi_gpu_checks(args, world, rank):
    """N > 1: (i) results on hardware -- a small slab of every element family is
    assembled in halo and in exchange mode (NCCL all_to_all) and compared on rank
    0 with the single-GPU assembly of the global mesh; (ii) the cost of the
    north star's interface exchange beside the halo mode on the headline strip."""
    import torch
    import torch.distributed as dist
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    from spfem_b200.distributed import (DistributedAssembler, slab_part, slab_global,
                                        strip_part)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    checked = []
    small = [("tri", 24, lambda: sp.ElementH1Vec(sp.ElementTriP2()), None, elasticity, vec_nitsche),
             ("tri", 24, lambda: sp.ElementH1Vec(sp.ElementTriP2()), sp.ElementTriP1, stokes_b, None),
             ("tri", 32, sp.ElementTriP1, None, lap2, None),
             ("tet", 6, sp.ElementTetP2, None, lap3, None)]
    for kind, size, feu, fev, form, facet in small:
        eu = feu()
        ev = fev() if fev is not None else None
        ref = fref = None
        if rank == 0:
            ga = AssemblerElement(slab_global(kind, size, world), eu, ev)
            ref = ga.iasm(form)
            fref = ga.fasm(facet) if facet else None
        for mode in ("halo", "exchange"):
            part = slab_part(kind, size, rank, world, eu, ev, halo=(mode == "halo"))
            da = DistributedAssembler(None, eu, ev, part=part, rank=rank, world=world, mode=mode)
            res = da.iasm(form).gather(0)
            fres = da.fasm(facet).gather(0) if facet else None
            if rank == 0:
                cases.compare_csr(res, ref)
                if facet:
                    cases.compare_csr(fres, fref)
                checked.append("%s/%s/%s" % (kind, form.__name__, mode))
    # exchange vs halo on the headline strip (ms per assembly through the API)
    timing = {}
    for mode in ("halo", "exchange"):
        part = strip_part(rank, world, args.refine, halo=(mode == "halo"))
        da = DistributedAssembler(None, sp.ElementTriP1(), part=part, rank=rank, world=world,
                                  mode=mode)
        for _ in range(3):
            res = da.iasm(lap2)
        torch.cuda.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            res = da.iasm(lap2)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        timing[mode + "_ms"] = float(tt.item())
        if mode == "exchange":
            sent = torch.tensor([res.xchg.interface_entries], device="cuda", dtype=torch.int64)
            dist.all_reduce(sent, op=dist.ReduceOp.MAX)
            timing["interface_entries_sent_max"] = int(sent.item())
        del da, res
        torch.cuda.empty_cache()
    return checked, timing


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import spfem_b200 as sp
    from spfem_b200.assembly import AssemblerElement
    t_start = time.perf_counter()

    # N = 1: the BASELINE configs[1] mesh itself.  N > 1 (weak scaling, fixed
    # work per GPU): a strip of N such unit-square meshes glued along x; rank r
    # assembles copy r plus the halo elements that touch rows it owns
    # (spfem_b200/distributed.py) -- no data-path collective.
    t0 = time.perf_counter()
    if world == 1:
        m = sp.MeshTri()
        m.refine(args.refine)
        a = AssemblerElement(m, sp.ElementTriP1())
        nt = m.t.shape[1]
    else:
        from spfem_b200.distributed import DistributedAssembler, strip_part
        part = strip_part(rank, world, args.refine)
        da = DistributedAssembler(None, sp.ElementTriP1(), part=part, rank=rank, world=world)
        a = da.local
        m = part.mesh
        nt = part.n_own_elements
    t_mesh = time.perf_counter() - t0
    nv = m.p.shape[1]
    t0 = time.perf_counter()
    A0 = a.iasm(FORMS[0][1])                    # first call through the public API
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    del A0
    t0 = time.perf_counter()
    preps = [a.prepare_iasm(f) for _, f in FORMS]
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    nnz = preps[0].pattern.nnz
    eng = a._get_engine()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident steps -------------------------------------------
    # the sampler thread starts before the warm-up (the first NVML queries of a
    # process can take tens of ms) and records only inside the timed region
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        for p in preps:
            p.launch()
    barrier()
    sampler.recording = True
    fused = all(p.n_launches == 1 for p in preps)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(1 + 2 * len(preps))]
          for _ in range(args.steps)]
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    start.record()
    if fused:
        # one kernel per assembly: nothing but the launches inside the timed region
        # (an event pair around every 0.1 ms launch costs 3-4 % of it); the kernel's
        # average launch duration over the region is region / number of launches
        for s in range(args.steps):
            for p in preps:
                p.launch()
    else:
        for s in range(args.steps):
            ev[s][0].record()
            for i, p in enumerate(preps):
                p.launch_element_kernel()
                ev[s][1 + 2 * i].record()
                p.launch_reduce()
                ev[s][2 + 2 * i].record()
    stop.record()
    barrier()
    clocks = sampler.stop()
    ms_total = start.elapsed_time(stop)
    if fused:
        t_elem = 0.0
        t_red = ms_total / (args.steps * len(preps))
    else:
        t_elem = np.mean([ev[s][2 * i].elapsed_time(ev[s][1 + 2 * i])
                          for s in range(args.steps) for i in range(len(preps))])
        t_red = np.mean([ev[s][1 + 2 * i].elapsed_time(ev[s][2 + 2 * i])
                         for s in range(args.steps) for i in range(len(preps))])
    # per form, outside the timed region: events around 50 back-to-back launches
    per_form = {}
    for (fname, _), p in zip(FORMS, preps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            p.launch()
        e1.record()
        torch.cuda.synchronize()
        per_form[fname + "_ms"] = e0.elapsed_time(e1) / 50
    if fused:
        kname = {"FanPrepared": "spfem_elem_fan", "TiledPrepared": "spfem_elem_tile"}.get(
            type(preps[0]).__name__, "fused")
        if kname == "spfem_elem_fan":
            kname += {"tiny": "8t", "fan8": "8"}.get(preps[0].fp.get("format"), "")
        kernels = {kname + "_ms": float(t_red)}
        t_elem = 0.0
    else:
        kernels = {"spfem_elem_ms": float(t_elem), "k_reduce_ms": float(t_red)}
    kernels["per_form"] = per_form
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    elems_per_step = 2 * nt * world
    value = elems_per_step / (ms_per_step * 1e-3)

    # ---- end to end through the public API (host buffers) ------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(max(args.warmup, 3)):        # untimed warm-up steps of the same loop
        eng.upload_mesh()
        for _, f in FORMS:
            A = a.iasm(f)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    t_up = 0.0
    for _ in range(e2e_steps):
        tu = time.perf_counter()
        eng.upload_mesh()                       # H2D of p and t (host inputs)
        t_up += time.perf_counter() - tu
        for _, f in FORMS:
            A = a.iasm(f)                       # scipy CSR on the host
            d2h += A.data.nbytes
    barrier()
    dt = time.perf_counter() - t0
    e2e_upload_ms = 1e3 * t_up / e2e_steps
    if world > 1:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    e2e_value = elems_per_step * e2e_steps / dt
    h2d_bytes = int(eng.h2d_bytes)
    tile_info = ({"kind": preps[0].tp.get("kind", "element-tile"),
                  "E": preps[0].tp["E"], "ldl": preps[0].tp.get("ldl"),
                  "ntiles": preps[0].tp["ntiles"],
                  "halo_factor": preps[0].tp["halo_factor"],
                  "tiles_per_cta": preps[0].tp.get("tiles_per_cta"),
                  "smem_bytes": preps[0].smem,
                  "threads": preps[0].threads} if fused else None)
    n_launch = int(args.steps * sum(p.n_launches for p in preps))
    del preps, A, a, eng
    torch.cuda.empty_cache()

    # ---- peaks ---------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic_all = {}
    for tname in ("traffic_r02.json", "traffic_r01.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            try:
                traffic_all = json.load(open(tpath))
                break
            except Exception:
                traffic_all = {}
    cfg_traffic = traffic_all.get("configs_per_element", {})

    # ---- N > 1: results on hardware + cost of the interface exchange -------------
    parity_checked, exchange = None, None
    if world > 1 and not args.no_parity:
        parity_checked, exchange = multi_gpu_checks(args, world, rank)

    # ---- the other configurations -------------------------------------------------
    cfgs = []
    names = [c for c in args.configs if c in CONFIGS]
    for name in names:
        if time.perf_counter() - t_start > args.time_budget:
            cfgs.append({"name": name, "skipped": "time budget of %d s used up" % args.time_budget})
            continue
        try:
            cfgs.append(run_config(name, args, world, rank, peak, cfg_traffic, barrier))
        except Exception as exc:                 # keep the headline line: report the failure
            import traceback
            cfgs.append({"name": name, "error": repr(exc)[:300],
                         "where": traceback.format_exc().strip().splitlines()[-3:]})
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the headline kernel -------------------------------------------
    balg = algorithmic_bytes(nt, nv, nnz)
    achieved = balg / ((t_elem + t_red) * 1e-3) / 1e9
    traffic = traffic_all.get("dram_bytes_per_assembly")

    # ---- CPU baseline beside it (bounded sample) ----------------------------
    cpu = None
    if not args.no_cpu_baseline:
        step, cnt, kind = cpu_assemble_factory(args.cpu_refine)
        step()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            step()
        cdt = time.perf_counter() - t0
        t0 = time.perf_counter()
        step(best_case=True)
        bdt = time.perf_counter() - t0
        cpu = {"value": 2.0 * cnt * reps / cdt, "unit": UNIT, "cores": cpu_threads(),
               "kind": kind, "value_tind_arange": 2.0 * cnt / bdt,
               "sample": "MeshTri().refine(%d): %d triangles x 2 forms x %d reps, "
                         "AssemblerElement.iasm as shipped" % (args.cpu_refine, cnt, reps)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("MeshTri.refine(%d) ElementTriP1 Laplace+mass iasm" % args.refine)
                               + ("" if world == 1 else
                                  " x %d (strip of unit squares, one per GPU, halo "
                                  "recompute, no data-path collective)" % world),
                   "nt_per_gpu": nt, "nv": nv, "nnz": nnz,
                   "elements_per_step": elems_per_step,
                   "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2"
                         % ((balg + 72 * nt) * 2 / 1e6),
                   "mesh_build_s": round(t_mesh, 2), "setup_s": round(t_setup, 2),
                   "first_call_s": round(t_first, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src,
                     "kernels": kernels, "fused": bool(fused),
                     "tile": tile_info,
                     "algorithmic_bytes_per_assembly": balg},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                "ms_per_step": 1e3 * dt / e2e_steps, "upload_ms": e2e_upload_ms,
                "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(d2h // e2e_steps)},
        "gpu_launches": n_launch,
        "clocks": clocks,
        "configs": cfgs,
    }
    if world > 1:
        line["parity_checked"] = bool(parity_checked)
        line["parity_cases"] = parity_checked
        line["exchange"] = exchange
        if exchange:
            line["exchange_ms"] = exchange.get("exchange_ms")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refine", type=int, default=11)
    ap.add_argument("--cpu-refine", type=int, default=8, dest="cpu_refine")
    ap.add_argument("--e2e-steps", type=int, default=10, dest="e2e_steps")
    ap.add_argument("--no-cpu-baseline", action="store_true", dest="no_cpu_baseline")
    ap.add_argument("--ref-refine", type=int, default=10, dest="ref_refine")
    ap.add_argument("--ref-seconds", type=float, default=60.0, dest="ref_seconds")
    ap.add_argument("--configs", default="all",
                    help="'all', 'none' or a comma separated list of " + ", ".join(CONFIGS))
    ap.add_argument("--config-steps", type=int, default=10, dest="config_steps")
    ap.add_argument("--size", action="append", default=[],
                    help="name=size overrides of a configuration's mesh size")
    ap.add_argument("--time-budget", type=int, default=420, dest="time_budget",
                    help="seconds after which the remaining configurations are skipped")
    ap.add_argument("--no-parity", action="store_true", dest="no_parity")
    args = ap.parse_args()
    args.configs = (list(CONFIGS) if args.configs == "all" else
                    [] if args.configs == "none" else args.configs.split(","))
    args.sizes = dict(kv.split("=") for kv in args.size)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
