# -*- coding: utf-8 -*-
"""Worker of tests/test_distributed.py: world_size-N run of the partitioned
assembly with the gloo backend on the CPU.  The rank-local assembly is the host
emulation of the generated kernels (tests/hostemu.py); everything else --
partitioning, ownership, halo, gather -- is the product code."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch.distributed as dist

import spfem_b200 as sp
from spfem_b200.distributed import (DistributedAssembler, strip_part, strip_global,
                                    slab_part, slab_global)
import cases
import hostemu
import util


def host_assemble(asm, form, intorder=None, interp=None):
    return hostemu.iasm(asm, form, intorder=intorder, interp=interp)


def host_assemble_facet(asm, form, find=None, interior=False, intorder=None,
                         normals=True, interp=None):
    return hostemu.fasm(asm, form, find=find, interior=interior, intorder=intorder,
                        normals=normals, interp=interp)


def check_row_slices(res, full, rank, world):
    """The device row slices (``csr()``) of all ranks stacked = the gathered matrix."""
    urows, indptr, cols, vals = res.csr()
    objs = [None] * world
    dist.all_gather_object(objs, (urows.numpy(), indptr.numpy(), cols.numpy(), vals.numpy()))
    if rank != 0:
        return
    seen = np.concatenate([o[0] for o in objs])
    assert len(np.unique(seen)) == len(seen)            # every row on one rank only
    full = full.tocsr()
    for ur, ip, cc, vv in objs:
        for k, r in enumerate(ur):
            lo, hi = full.indptr[r], full.indptr[r + 1]
            assert np.array_equal(cc[ip[k]:ip[k + 1]], full.indices[lo:hi])
            assert np.array_equal(vv[ip[k]:ip[k + 1]], full.data[lo:hi])
    assert len(seen) == np.count_nonzero(np.diff(full.indptr))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    # 1. replicated global meshes, general partitioner
    for name in ("trip1_lap", "trip2_mass", "vecp2_elasticity", "stokes_b",
                 "tetp2_lap", "trip1_load_sin_h", "q2_lap"):
        case = cases.BY_NAME[name]
        m = util.mesh_for(case)
        eu = util.element(case.eu, case.ncu)
        ev = None
        if not (case.ev is None and case.ncv == case.ncu):
            ev = util.element(case.ev or case.eu, case.ncv)
        da = DistributedAssembler(m, eu, ev, assemble=host_assemble)
        dres = da.iasm(case.form)
        res = dres.gather(0)
        if dres.bilinear:
            check_row_slices(dres, res, rank, world)
        if rank == 0:
            util.check(res, util.golden(case))
            print("dist ok", name, "world", world)
    # 1b. the same through the interface exchange (own block only, ghost rows
    #     summed on the owner in fixed rank order over one all_to_all)
    for name in ("trip1_lap", "trip2_mass", "vecp2_elasticity", "stokes_b",
                 "tetp2_lap", "trip1_load_sin_h", "q2_lap"):
        case = cases.BY_NAME[name]
        m = util.mesh_for(case)
        eu = util.element(case.eu, case.ncu)
        ev = None
        if not (case.ev is None and case.ncv == case.ncu):
            ev = util.element(case.ev or case.eu, case.ncv)
        da = DistributedAssembler(m, eu, ev, assemble=host_assemble, mode="exchange")
        r1 = da.iasm(case.form)
        r2 = da.iasm(case.form)                   # reuses the exchange plan
        assert len(da._xchg) == 1
        assert np.array_equal(r1.values.numpy(), r2.values.numpy())   # reproducible
        assert r1.xchg.interface_entries > 0 or world == 1 or rank == 0
        # every rank holds exactly the rows it owns
        rows = np.unique(r1.xchg.rows.numpy())
        assert np.array_equal(rows, np.sort(da.part.l2g_v[da.part.owned_v])), name
        urows, indptr, cols, vals = r1.csr()
        assert int(indptr[-1]) == vals.numel() == cols.numel()
        res = r1.gather(0)
        if r1.bilinear:
            check_row_slices(r1, res, rank, world)
        if rank == 0:
            util.check(res, util.golden(case))
            print("xchg ok", name, "world", world)
    # 1c. boundary-facet forms (the Nitsche terms of config 5) in both modes:
    #     only the global boundary is assembled, not the cuts
    for name in ("f_trip1_nitsche", "f_trip1_lin", "f_trip2_nitsche", "f_vecp2_nitsche",
                 "f_tetp2_bnd", "f_q1_bnd"):
        case = cases.BY_NAME[name]
        m = util.mesh_for(case)
        eu = util.element(case.eu, case.ncu)
        for mode in ("halo", "exchange"):
            da = DistributedAssembler(m, eu, None, assemble=host_assemble,
                                      assemble_facet=host_assemble_facet, mode=mode)
            nloc = len(da.part.global_boundary_facets())
            assert nloc <= len(m.boundary_facets())
            assert mode == "halo" or nloc < len(m.boundary_facets())
            res = da.fasm(case.form, normals=case.kw.get("normals", True)).gather(0)
            if rank == 0:
                util.check(res, util.golden(case))
                print("facet ok", name, mode, "world", world)
    # 1d. interior-facet (DG / jump) forms: facet-neighbour halo
    for name in ("f_trip1_sipg", "f_tridg1_sipg", "f_tetp2_jump"):
        case = cases.BY_NAME[name]
        m = util.mesh_for(case)
        eu = util.element(case.eu, case.ncu)
        da = DistributedAssembler(m, eu, None, assemble=host_assemble,
                                  assemble_facet=host_assemble_facet, facet_halo=True)
        res = da.fasm(case.form, interior=True).gather(0)
        ia = da.iasm(cases.mass).gather(0)          # interior forms on the wider halo
        if rank == 0:
            util.check(res, util.golden(case))
            print("interior facet ok", name, "world", world)
        try:
            DistributedAssembler(m, eu, None, assemble_facet=host_assemble_facet).fasm(
                case.form, interior=True)
            raise AssertionError("interior facets without facet_halo must be refused")
        except NotImplementedError:
            pass
    # 2. weak-scaling strip mesh built per rank, against the oracle on the
    #    explicitly assembled global strip
    part = strip_part(rank, world, 3)
    da = DistributedAssembler(None, sp.ElementTriP1(), part=part, assemble=host_assemble)
    for form in (cases.lap2, cases.mass, cases.load1):
        res = da.iasm(form).gather(0)
        if rank == 0:
            import asm_oracle
            g = strip_global(world, 3)
            ref = asm_oracle.iasm(g, "TriP1", form)
            if hasattr(ref, "indptr"):
                cases.compare_csr(res, ref)
            else:
                cases.compare_vec(res, ref)
            print("strip ok", form.__name__)
    # 2b. the strip without halo, through the interface exchange
    part = strip_part(rank, world, 3, halo=False)
    assert part.mesh.t.shape[1] == part.n_own_elements
    da = DistributedAssembler(None, sp.ElementTriP1(), part=part, assemble=host_assemble,
                              mode="exchange")
    for form in (cases.lap2, cases.mass, cases.load1):
        res = da.iasm(form).gather(0)
        if rank == 0:
            import asm_oracle
            ref = asm_oracle.iasm(strip_global(world, 3), "TriP1", form)
            if hasattr(ref, "indptr"):
                cases.compare_csr(res, ref)
            else:
                cases.compare_vec(res, ref)
            print("strip exchange ok", form.__name__)
    # 2c. boundary forms on the strip: the cuts x = 1 .. world-1 are not boundary
    for mode in ("halo", "exchange"):
        part = strip_part(rank, world, 3, halo=(mode == "halo"))
        da = DistributedAssembler(None, sp.ElementTriP1(), part=part, mode=mode,
                                  assemble_facet=host_assemble_facet)
        for form in (cases.nitsche, cases.bnd_lin):
            res = da.fasm(form).gather(0)
            if rank == 0:
                import asm_oracle
                ref = asm_oracle.fasm(strip_global(world, 3), "TriP1", form)
                util.check(res, ref)
                print("strip facet ok", mode, form.__name__)
    # 3. slab partition of the structured benchmark grids (configs 3-5): any
    #    element, rank-local mesh generation, both modes, interior + boundary forms
    slab_cases = [("tri", 4, sp.ElementH1Vec(sp.ElementTriP2()), None, cases.elasticity, "f"),
                  ("tri", 4, sp.ElementH1Vec(sp.ElementTriP2()), sp.ElementTriP1(), cases.stokes_b, None),
                  ("tri", 5, sp.ElementTriP1(), None, cases.mass, None),
                  ("tet", 2, sp.ElementTetP2(), None, cases.lap3, None)]
    for kind, size, eu, ev, form, facet in slab_cases:
        gm = slab_global(kind, size, world)
        ref_asm = sp.AssemblerElement(gm, eu, ev)
        ref = hostemu.iasm(ref_asm, form)
        for mode in ("halo", "exchange"):
            part = slab_part(kind, size, rank, world, eu, ev, halo=(mode == "halo"))
            da = DistributedAssembler(None, eu, ev, part=part, assemble=host_assemble,
                                      assemble_facet=host_assemble_facet, mode=mode)
            res = da.iasm(form).gather(0)
            if rank == 0:
                cases.compare_csr(res, ref)
                print("slab ok", kind, mode, form.__name__)
            if facet:
                fres = da.fasm(cases.vec_nitsche).gather(0)
                if rank == 0:
                    cases.compare_csr(fres, hostemu.fasm(ref_asm, cases.vec_nitsche))
                    print("slab facet ok", kind, mode)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
