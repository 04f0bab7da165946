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
    refine = args.cpu_refine
    step, nt, kind = cpu_assemble_factory(refine)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = 2.0 * nt * args.steps / dt
    sample = ("MeshTri().refine(%d): %d triangles x 2 forms per step, %s "
              "AssemblerElement.iasm as shipped (tind=None)" % (refine, nt,
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
    for s in range(args.steps):
        ev[s][0].record()
        for i, p in enumerate(preps):
            if fused:
                ev[s][1 + 2 * i].record()
                p.launch()
            else:
                p.launch_element_kernel()
                ev[s][1 + 2 * i].record()
                p.launch_reduce()
            ev[s][2 + 2 * i].record()
    stop.record()
    barrier()
    clocks = sampler.stop()
    ms_total = start.elapsed_time(stop)
    t_elem = np.mean([ev[s][2 * i].elapsed_time(ev[s][1 + 2 * i])
                      for s in range(args.steps) for i in range(len(preps))])
    t_red = np.mean([ev[s][1 + 2 * i].elapsed_time(ev[s][2 + 2 * i])
                     for s in range(args.steps) for i in range(len(preps))])
    if fused:
        kname = {"FanPrepared": "spfem_elem_fan", "TiledPrepared": "spfem_elem_tile"}.get(
            type(preps[0]).__name__, "fused")
        if kname == "spfem_elem_fan":
            kname += {"tiny": "8t", "fan8": "8"}.get(preps[0].fp.get("format"), "")
        kernels = {kname + "_ms": float(t_red)}
        t_elem = 0.0
    else:
        kernels = {"spfem_elem_ms": float(t_elem), "k_reduce_ms": float(t_red)}
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the two kernels that make up one assembly --------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    balg = algorithmic_bytes(nt, nv, nnz)
    achieved = balg / ((t_elem + t_red) * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_assembly")
        except Exception:
            traffic = None

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
                   "mesh_build_s": round(t_mesh, 2), "setup_s": round(t_setup, 2)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src,
                     "kernels": kernels, "fused": bool(fused),
                     "tile": ({"kind": preps[0].tp.get("kind", "element-tile"),
                               "E": preps[0].tp["E"], "ldl": preps[0].tp["ldl"],
                               "ntiles": preps[0].tp["ntiles"],
                               "halo_factor": preps[0].tp["halo_factor"],
                               "smem_bytes": preps[0].smem,
                               "threads": preps[0].threads} if fused else None),
                     "algorithmic_bytes_per_assembly": balg},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                "ms_per_step": 1e3 * dt / e2e_steps, "upload_ms": e2e_upload_ms,
                "h2d_bytes_per_step": int(eng.h2d_bytes),
                "d2h_bytes_per_step": int(d2h // e2e_steps)},
        "gpu_launches": int(args.steps * sum(p.n_launches for p in preps)),
        "clocks": clocks,
    }
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
