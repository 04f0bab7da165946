# -*- coding: utf-8 -*-
"""CUDA C generation for the element / facet kernels.

Input: the traced normal form of a weak form (``tracer.Form`` per pair of
vector components) plus the reference element tables.  Output: the source of
one ``extern "C" __global__`` kernel for sm_100a, compiled by NVRTC through
the C-ABI (``spfem_nvrtc_compile``).

Interior kernels (``AssemblerElement.iasm``, spfem/assembly.py:919-1006)
----------------------------------------------------------------------
One thread per element.  The thread loads the element's vertex ids and
coordinates, rebuilds the affine geometry of ``MappingAffine``
(spfem/mapping.py:207-302; same operation order, ``__dmul_rn/__dadd_rn`` so no
FMA contraction changes ``x`` -- forms compare coordinates with ``==``) and
evaluates the few coefficient expressions ``G^{ab}`` of the traced form.  Every
local matrix entry is then a short dot product

    A_loc[j, i] = sum_terms  K_term(j, i) * G_term

with *compile-time* constants ``K``: for element-constant coefficients the sum
over quadrature points is folded on the host into the reference tensors
``K^{ab}_{ji} = sum_q W_q psi^a_j(X_q) psi^b_i(X_q)``; coefficients that vary
from point to point (``x``, ``w``, bilinear ``MappingQ1`` Jacobians) keep one
term per quadrature point.  Entries whose constants are all zero are known at
generation time and cost nothing.

Facet kernels (``fasm``, spfem/assembly.py:1011-1183) evaluate the basis
polynomials at run time at the per-facet reference points
``Y = invF(G(X))`` and loop over points and basis pairs; there are O(sqrt(nt))
boundary facets, so that kernel is written for generality, not speed.

The element body is a ``SPFEM_DEVICE`` function so the identical text can be
compiled for the host (``-DSPFEM_HOST``, tests only) to check generated code
against the oracle without a GPU.
"""
import hashlib
import os
import numpy as np
from .tracer import NONE

PRELUDE = r"""
#ifdef SPFEM_HOST
#include <math.h>
#define SPFEM_DEVICE static inline
#define __dmul_rn(a, b) ((a) * (b))
#define __dadd_rn(a, b) ((a) + (b))
#define __dsub_rn(a, b) ((a) - (b))
#define __ddiv_rn(a, b) ((a) / (b))
#define __restrict__
struct double2 { double x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
#else
#define SPFEM_DEVICE static __device__ __forceinline__
#endif
#if defined(SPFEM_HOST) || defined(SPFEM_IEEE_RCP)
#define spfem_rcp(x) (1.0 / (x))
#else
/* 1/x for the Jacobian determinant: MUFU.RCP64H seed (20 bits) + two Newton steps
   (5 instructions, <= 1 ulp from the IEEE quotient for normal x; the compiler's
   1.0/x adds a correction step and a range check with a slow-path call).      */
static __device__ __forceinline__ double spfem_rcp(const double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
#endif
#define SPFEM_MAX_PARAMS 32
struct SpfemArgs {
    long long n;              /* number of elements / facets to process        */
    long long ldo;            /* leading dimension of the entry-major output    */
    const int *sel;           /* optional element subset (tind) or null         */
    const int *t;             /* vertex connectivity, t[k*ldt + e]              */
    long long ldt;
    const double *p;          /* vertex coordinates, p[k*ldp + v]               */
    long long ldp;
    const int *tdof_u;        /* trial DOF table (interp only), [j*ldd + e]     */
    long long ldd;
    const double *interp[8];  /* interpolated solution vectors (w)              */
    const int *fv;            /* facet kernels: facet vertex ids fv[k*n + f]    */
    const int *e1;            /* facet kernels: first / second element          */
    const int *e2;
    const int *lf1;           /* facet kernels: local facet index inside e1     */
    double *out;              /* out[entry*ldo + k*ldk]                         */
    long long ldk;
    double params[SPFEM_MAX_PARAMS]; /* captured Python scalars (dt, penalties ...) */
    const double *fields[8];  /* host-evaluated coefficient fields, fields[i][q*ldf + e] */
    long long ldf;
};
"""

# Fused tile kernel: one CTA per tile of (Morton-contiguous) elements plus the
# halo elements that touch rows the tile owns (tileplan.py; plan built natively by
# csrc/spfem_plan.cu).  The per-tile blob (vertex ids, coordinates, contribution
# lists, CSR positions) is brought into shared memory by two TMA bulk copies
# (cp.async.bulk + mbarrier, SASS UBLKCP).  Phase A: every thread computes one
# local matrix into shared memory (distinct values only, sL[u*ldl + el]).  Phase
# B: one thread per slot owned by the tile sums its contributions from shared
# memory in a fixed order and writes the CSR value(s) exactly once -- no atomics,
# no partial sums in HBM.  For vector elements (SPFEM_NB = bsv*bsu > 1) a slot
# is the bsv x bsu block of values of one scalar entry: one index decode serves
# NB values and the plan is NB times smaller.
TILE_THREADS = int(os.environ.get("SPFEM_TILE_T", "256"))
DEVICE_HELPERS = r"""
#ifndef SPFEM_HOST
static __device__ __forceinline__ unsigned spfem_smem(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
static __device__ __forceinline__ void spfem_mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(spfem_smem(bar)), "r"(count));
}
static __device__ __forceinline__ void spfem_bulk_load(void *dst, const void *src,
                                                       unsigned bytes, void *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(spfem_smem(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(spfem_smem(dst)), "l"(src), "r"(bytes), "r"(spfem_smem(bar)) : "memory");
}
static __device__ __forceinline__ void spfem_mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SPFEM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SPFEM_DONE;\n"
        "bra SPFEM_WAIT;\n"
        "SPFEM_DONE:\n"
        "}\n" ::"r"(spfem_smem(bar)), "r"(parity) : "memory");
}
#endif
"""
TILE_KERNEL = r"""
struct SpfemTileArgs {
    SpfemArgs base;
    const int *desc;              /* 16 ints per tile (tileplan.py)              */
    const unsigned char *blob;    /* packed per-tile data, 16-byte aligned       */
    double *values;               /* CSR values                                   */
    int ldl;                      /* leading dimension of the smem tile           */
    int max_a, max_b;             /* smem bytes reserved for blob sections A / B  */
    int compact;                  /* packed local values (per-element bit masks)  */
    int store;                    /* doubles reserved for the local values        */
    int prefetch;                 /* L2 prefetch distance in tiles (0 = off)      */
    int sorted;                   /* tile slots ordered by contribution count:
                                     position inside the row comes from `spos`    */
    int els;                      /* block plans: contribution = entry << els | element */
};
/* Packed shared-memory image of one element's local values (compact plans):
   only the values whose mask bit is set are stored, value u at
   base[number of set bits below u].  The element body writes through
   out[u] = val, u a literal after inlining.                                   */
struct SpfemPacked {
    double *base;
    unsigned m[SPFEM_NMASK];
    int pre[SPFEM_NMASK];
    struct Ref {
        double *p; bool on;
        __device__ __forceinline__ void operator=(const double v) const { if (on) *p = v; }
    };
    __device__ __forceinline__ Ref operator[](const long long u) const {
        const int w = (int)(u >> 5), b = (int)(u & 31);
        Ref r;
        r.on = ((m[w] >> b) & 1u) != 0u;
        r.p = base + pre[w] + __popc(m[w] & ((1u << b) - 1u));
        return r;
    }
};
#if SPFEM_NB > 1
/* distinct-value id of component (a, b) of scalar entry j*nbv + i (block plans) */
__device__ const unsigned short SPFEM_UT[SPFEM_NSIDX * SPFEM_NB] = {@UT@};
/* one row of the (id * ldl) table in shared memory -> registers, as one vector
   load where the row is 8 or 16 bytes (the table starts 16-byte aligned)     */
#if SPFEM_NB == 4
#define SPFEM_UT_LOAD(u, tab, r) do { const int4 q_ = reinterpret_cast<const int4 *>(tab)[r]; \
    u[0] = q_.x; u[1] = q_.y; u[2] = q_.z; u[3] = q_.w; } while (0)
#elif SPFEM_NB == 2
#define SPFEM_UT_LOAD(u, tab, r) do { const int2 q_ = reinterpret_cast<const int2 *>(tab)[r]; \
    u[0] = q_.x; u[1] = q_.y; } while (0)
#else
#define SPFEM_UT_LOAD(u, tab, r) do { _Pragma("unroll") \
    for (int z_ = 0; z_ < SPFEM_NB; ++z_) u[z_] = (tab)[(r) * SPFEM_NB + z_]; } while (0)
#endif
#endif
extern "C" __global__ void __launch_bounds__(@THREADS@, @MINB@) @NAME@_tile(const SpfemTileArgs T)
{
    extern __shared__ __align__(128) unsigned char spfem_smem_raw[];
    const int ldl = T.ldl;
    double *sL = (double *)spfem_smem_raw;
    unsigned char *sA = spfem_smem_raw + (((size_t)8 * T.store + 15) & ~(size_t)15);
    unsigned char *sB = sA + T.max_a;
    unsigned long long *bars = (unsigned long long *)(sB + T.max_b);
    const int *d = T.desc + 16 * (size_t)blockIdx.x;
    const int bytes_a = d[1], bytes_b = d[2], ne = d[3], ne_pad = d[4], ns = d[5];
    if (ne == 0 || ns == 0) return;
#if SPFEM_NB > 1
    int *sUT = (int *)(bars + 2);             /* id * ldl, one row of NB ints per entry */
    for (int i = threadIdx.x; i < SPFEM_NSIDX * SPFEM_NB; i += blockDim.x)
        sUT[i] = (int)SPFEM_UT[i] * ldl;
#endif
    if (threadIdx.x == 0) {
        spfem_mbar_init(&bars[0], 1);
        spfem_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned char *src = T.blob + (size_t)(unsigned)d[0] * 16;
        spfem_bulk_load(sA, src, (unsigned)bytes_a, &bars[0]);
        spfem_bulk_load(sB, src + bytes_a, (unsigned)bytes_b, &bars[1]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        /* pull the blob of a tile that starts a few waves later into L2 (after the
           barrier: the descriptor load of that tile must not hold up the CTA)   */
        const long long tp = (long long)blockIdx.x + T.prefetch;
        if (T.prefetch > 0 && tp < (long long)gridDim.x) {
            const int *dp = T.desc + 16 * (size_t)tp;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                         ::"l"(T.blob + (size_t)(unsigned)dp[0] * 16),
                           "r"((unsigned)(dp[1] + dp[2])) : "memory");
        }
    }
    /* phase A: local matrices -> sL */
    spfem_mbar_wait(&bars[0], 0);
    const unsigned short *vt = (const unsigned short *)sA;   /* tile-local vertex ids */
    const double *pc = (const double *)(sA + d[12]);          /* tile-local coordinates */
    const long long ldpc = d[11];
    /* SPFEM_NPART threads per element, each computing every NPART-th distinct
       local value (large local matrices: keeps all threads of the CTA busy); the
       part index is uniform inside a warp                                       */
    const int ne32 = (ne + 31) & ~31;
    if (T.compact) {
        const int *sbase = (const int *)(sA + d[15]);
        const unsigned *smask = (const unsigned *)(sbase + ne_pad);
        for (int w = threadIdx.x; w < ne32 * SPFEM_NPART; w += blockDim.x) {
            const int part = w / ne32, el = w - part * ne32;
            if (el >= ne) continue;
            SpfemPacked po;
            po.base = sL + sbase[el];
            int below = 0;
#pragma unroll
            for (int q = 0; q < SPFEM_NMASK; ++q) {
                po.m[q] = smask[q * ne_pad + el];
                po.pre[q] = below;
                below += __popc(po.m[q]);
            }
            switch (part) {
@PCASES@
            }
        }
    } else {
        for (int w = threadIdx.x; w < ne32 * SPFEM_NPART; w += blockDim.x) {
            const int part = w / ne32, el = w - part * ne32;
            if (el >= ne) continue;
            switch (part) {
@CASES@
            }
        }
    }
    /* phase B: one thread per slot of the tile's rows (a slot = one CSR value, or
       the bsv x bsu block of values of a vector element's scalar entry); slots in
       CSR order, so a warp stores runs of consecutive values.  Contributions are
       added in list order => bitwise reproducible.                              */
    spfem_mbar_wait(&bars[1], 0);
    __syncthreads();
    /* Where a slot's (first) CSR value goes, as a 32-bit offset from the tile's lowest
       position.  Slots ordered by contribution count carry it themselves (`spos`, and
       `slen` = length of the scalar row for block plans).  Slots in CSR order: one
       header per 32 slots (bit l = slot l starts a new row, row of slot 0) and one
       record per row (offset, first slot, length) -- a quarter byte of plan per slot
       instead of 4 or 6, and the kernels are bound by HBM traffic.                 */
    const unsigned *spos = (const unsigned *)(sB + 16);
    const uint2 *rrec = (const uint2 *)(sB + 16);
    const uint2 *shdr = (const uint2 *)(sB + d[7]);
#define SPFEM_SLOT_PLACE(s, pos, L) \
    unsigned pos, L; \
    if (T.sorted) { pos = spos[s]; L = SPFEM_NB > 1 ? (unsigned)slen[s] : 0u; } \
    else { \
        const uint2 h_ = shdr[(s) >> 5]; \
        const uint2 r_ = rrec[h_.y + __popc(h_.x & (0xffffffffu >> (31 - ((s) & 31))))]; \
        pos = r_.x + SPFEM_BSU * ((unsigned)(s) - (r_.y & 0xffffu)); L = r_.y >> 16; \
    }
    const unsigned short *soff = (const unsigned short *)(sB + d[6]);
    const unsigned short *cb = (const unsigned short *)(sB + d[8]);
    double *__restrict__ vals = T.values + *reinterpret_cast<const long long *>(sB);
    const unsigned short *slen = (const unsigned short *)(sB + d[7]);
#if SPFEM_NB == 1
#pragma unroll @BUNROLL@
    for (int s = threadIdx.x; s < ns; s += blockDim.x) {
        const unsigned k0 = soff[s], k1 = soff[s + 1];
        const unsigned c0 = cb[k0], c1 = cb[k0 + 1];
        SPFEM_SLOT_PLACE(s, pos, L_);
        double acc = sL[c0];
        if (k0 + 1 < k1) acc += sL[c1];
#pragma unroll @TUNROLL@
        for (unsigned kk = k0 + 2; kk < k1; ++kk) acc += sL[cb[kk]];
        vals[pos] = acc;
    }
#else
    const unsigned elm = (1u << T.els) - 1u;
    for (int s = threadIdx.x; s < ns; s += blockDim.x) {
        const unsigned k0 = soff[s], k1 = soff[s + 1];
        SPFEM_SLOT_PLACE(s, pos, L);
        double acc[SPFEM_NB];
        {
            const unsigned c = cb[k0];
            int ut[SPFEM_NB];
            SPFEM_UT_LOAD(ut, sUT, c >> T.els);
            const double *col = sL + (c & elm);
#pragma unroll
            for (int ab = 0; ab < SPFEM_NB; ++ab) acc[ab] = col[ut[ab]];
        }
#pragma unroll 2
        for (unsigned kk = k0 + 1; kk < k1; ++kk) {
            const unsigned c = cb[kk];
            int ut[SPFEM_NB];
            SPFEM_UT_LOAD(ut, sUT, c >> T.els);
            const double *col = sL + (c & elm);
#pragma unroll
            for (int ab = 0; ab < SPFEM_NB; ++ab) acc[ab] += col[ut[ab]];
        }
        double *__restrict__ o = vals + pos;
#pragma unroll
        for (int a = 0; a < SPFEM_BSV; ++a) {
#if SPFEM_BSU == 2
            *reinterpret_cast<double2 *>(o) = make_double2(acc[2 * a], acc[2 * a + 1]);
#else
#pragma unroll
            for (int b = 0; b < SPFEM_BSU; ++b) o[b] = acc[a * SPFEM_BSU + b];
#endif
            o += SPFEM_BSU * L;
        }
    }
#endif
}
"""

# Vertex-fan kernel (scalar P1 on triangles, element-constant coefficients).
# One thread per CSR row = mesh vertex r.  The thread walks the triangles around
# r in counter-clockwise ring order, recomputes the local row of r for each of
# them from the vertex coordinates (registers; each triangle is computed by its
# three vertices -- redundant flops instead of shared-memory traffic) and keeps
# one accumulator per ring neighbour: slot (r, n_q) receives exactly the two
# triangles that share the edge, in a fixed order.  No local matrices in shared
# memory, no contribution lists, no atomics.  Per-row record (fanplan.py):
#   plane 0 {g0 - s0, s0 | nfan << 16 | closed << 24, neighbours 0..3}, plane 1
#   further ring neighbours (16-bit tile-local vertex ids); plus one byte per
#   slot mapping CSR order to the ring order of the staging buffer
FAN_THREADS = int(os.environ.get("SPFEM_FAN_T", "128")) // 32 * 32
FAN_KERNEL = r"""
struct SpfemFanArgs {
    SpfemArgs base;
    const int *desc;              /* 8 ints per tile: off16, bytes A, bytes B, nr, ns, nvt */
    const unsigned char *blob;
    double *values;
    int max_a, max_b, max_ns, prefetch;
};
template <int MAXN, bool TINY>
SPFEM_DEVICE void @NAME@_fan_row(const SpfemArgs &P, const double2 *__restrict__ pc2,
                                 const unsigned char *__restrict__ rec, const int nr,
                                 const int row, double *__restrict__ sV, int *__restrict__ sG)
{
    /* row record planes -> registers; neighbour ids (16 bit; TINY: 8 bit) and CSR
       positions (8 bit; TINY: 4 bit) are extracted with shifts                  */
    unsigned w[8], pw[3];
    {
        const uint4 h = reinterpret_cast<const uint4 *>(rec)[row];
        w[0] = h.x; w[1] = h.y; w[2] = h.z; w[3] = h.w;
        if (TINY) {
            w[4] = w[5] = w[6] = w[7] = 0u;
            pw[0] = reinterpret_cast<const unsigned *>(rec + (size_t)16 * nr)[row];
            pw[1] = pw[2] = 0u;
        } else {
        const uint4 n1 = reinterpret_cast<const uint4 *>(rec + (size_t)16 * nr)[row];
        if (MAXN <= 8) {
            w[4] = n1.x; w[5] = n1.y; w[6] = 0u; w[7] = 0u;
            pw[0] = n1.z; pw[1] = n1.w; pw[2] = 0u;
        } else {
            const uint4 n2 = reinterpret_cast<const uint4 *>(rec + (size_t)32 * nr)[row];
            w[4] = n1.x; w[5] = n1.y; w[6] = n1.z; w[7] = n1.w;
            pw[0] = n2.x; pw[1] = n2.y; pw[2] = n2.z;
        }
        }
    }
#define SPFEM_NBR(q) (TINY ? ((w[2 + (((q) & 7) >> 2)] >> (8 * ((q) & 3))) & 0xffu) \
                           : ((w[2 + ((q) >> 1)] >> (16 * ((q) & 1))) & 0xffffu))
#define SPFEM_POS(q) (TINY ? ((pw[0] >> (4 * ((q) & 7))) & 0xfu) \
                           : ((pw[(q) >> 2] >> (8 * ((q) & 3))) & 0xffu))
    const int g0 = (int)w[0];                    /* CSR position of the row */
    const int s0 = (int)(w[1] & 0xffffu);
    const int nfan = (int)((w[1] >> 16) & 0xfu);
    const bool closed = ((w[1] >> 20) & 0x1u) != 0u;
    const int posdiag = (int)((w[1] >> 24) & 0xffu);
    const double2 cr = pc2[row];                 /* rows are the first vertices of a tile */
    const int nn = nfan + (closed ? 0 : 1);      /* ring neighbours */
    /* Slot (r, n_f) receives exactly two triangles: f-1 (where n_f is the far
       neighbour) and f.  `carry` hands the first over to the next step, so the
       sum is carry + L[r, n_f] in ring order -- no accumulator array.  Slot 0 of a
       closed ring gets its second contribution from the last triangle.          */
    double diag = 0.0, carry = 0.0, a0 = 0.0;
    const double2 c0 = pc2[SPFEM_NBR(0)];
    double2 ca = c0;
#pragma unroll
    for (int f = 0; f < MAXN; ++f) {
        if (f >= nfan) break;
        double2 cb = c0;                         /* closed ring wraps to neighbour 0 */
        if (f + 1 < nn) cb = pc2[SPFEM_NBR((f + 1) < MAXN ? (f + 1) : MAXN - 1)];
        double L[SPFEM_NUNIQUE];
#ifdef SPFEM_FAN_NOCOMPUTE     /* experiment: same memory behaviour, no geometry / form */
#pragma unroll
        for (int z = 0; z < SPFEM_NUNIQUE; ++z) L[z] = cr.x + ca.y + cb.x;
#else
        @NAME@_core(P, cr.x, cr.y, ca.x, ca.y, cb.x, cb.y, L);
#endif
        diag += L[@A00@];                        /* (test r, trial r)       */
        if (f == 0) {
            a0 = L[@A10@];                       /* (test r, trial n_0), first half */
        } else {
            /* staged in RING order (slot s0 = diagonal, s0 + 1 + q = neighbour q):
               the lanes of a warp write with a fixed stride, free of bank conflicts;
               next to every value goes its CSR position                          */
            sV[s0 + 1 + f] = carry + L[@A10@];   /* (test r, trial n_f)     */
            sG[s0 + 1 + f] = g0 + (int)SPFEM_POS(f);
        }
        carry = L[@A20@];                        /* (test r, trial n_{f+1}) */
        ca = cb;
    }
    if (closed) {
        a0 += carry;                             /* the last triangle's n_{f+1} is n_0 */
    } else if (TINY) {                           /* open ring: the last neighbour's slot */
        sV[s0 + 1 + nfan] = carry;
        sG[s0 + 1 + nfan] = g0 + (int)((pw[0] >> (4 * (nfan & 7))) & 0xfu);
    } else {
#pragma unroll
        for (int q = 1; q < MAXN + 1; ++q)
            if (q == nfan) { sV[s0 + 1 + q] = carry; sG[s0 + 1 + q] = g0 + (int)SPFEM_POS(q < MAXN ? q : MAXN - 1); }
    }
    sV[s0] = diag;
    sG[s0] = g0 + posdiag;
    sV[s0 + 1] = a0;
    sG[s0 + 1] = g0 + (int)SPFEM_POS(0);
#undef SPFEM_POS
#undef SPFEM_NBR
}
#ifndef SPFEM_HOST
#if defined(SPFEM_FAN_STCS)
#define SPFEM_FAN_ST(p, v) __stcs((p), (v))
#elif defined(SPFEM_FAN_STCG)
#define SPFEM_FAN_ST(p, v) __stcg((p), (v))
#elif defined(SPFEM_FAN_STWT)
#define SPFEM_FAN_ST(p, v) __stwt((p), (v))
#else
#define SPFEM_FAN_ST(p, v) (*(p) = (v))
#endif
template <int MAXN, bool TINY>
static __device__ __forceinline__ void @NAME@_fan_tile(const SpfemFanArgs &T)
{
    extern __shared__ __align__(128) unsigned char spfem_smem_raw[];
    unsigned char *sA = spfem_smem_raw;                       /* the tile's blob */
    double *sV = (double *)(sA + T.max_a);
    unsigned long long *bars = (unsigned long long *)(sV + T.max_ns);
    int *sG = (int *)(bars + 2);
#ifdef SPFEM_FAN_X_SAMEBLOB   /* experiment: every CTA works on one of 64 L2-resident tiles */
    const int *d = T.desc + 8 * (size_t)(blockIdx.x & 63);
#else
    const int *d = T.desc + 8 * (size_t)blockIdx.x;
#endif
    const int bytes_a = d[1], bytes_b = d[2], nr = d[3], ns = d[4];
    if (nr == 0) return;
    if (threadIdx.x == 0) {
        spfem_mbar_init(&bars[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        spfem_bulk_load(sA, T.blob + (size_t)(unsigned)d[0] * 16,
                        (unsigned)(bytes_a + bytes_b), &bars[0]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        /* pull the blob of a tile that will start a few waves later into L2, so
           its own TMA load does not pay the DRAM latency (issued after the
           barrier: the descriptor load of that tile must not hold up this CTA)  */
        const long long tp = (long long)blockIdx.x + T.prefetch;
        if (T.prefetch > 0 && tp < (long long)gridDim.x) {
            const int *dp = T.desc + 8 * (size_t)tp;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                         ::"l"(T.blob + (size_t)(unsigned)dp[0] * 16),
                           "r"((unsigned)(dp[1] + dp[2])) : "memory");
        }
    }
    spfem_mbar_wait(&bars[0], 0);
    for (int row = threadIdx.x; row < nr; row += blockDim.x)
        @NAME@_fan_row<MAXN, TINY>(T.base, (const double2 *)sA, sA + bytes_a, nr, row, sV, sG);
#if defined(SPFEM_FAN_X_NOSTORE)          /* experiment: load + compute only */
    __syncthreads();
    for (int s = threadIdx.x; s < ns; s += blockDim.x) { const double v = sV[s]; if (v == 1.2345e-300) T.values[sG[s]] = v; }
#elif defined(SPFEM_FAN_X_LINEAR)         /* experiment: contiguous (wrong) store positions */
    __syncthreads();
    for (int s = threadIdx.x; s < ns; s += blockDim.x) T.values[d[6] + s] = sV[s];
#elif !defined(SPFEM_FAN_CTASTORE)
    /* the slots of a warp's 32 rows are one contiguous range of the staging
       buffer: no CTA-wide barrier, every warp stores as soon as it is done
       (blockDim.x is a multiple of 32)                                          */
    {
        const int wrow = (int)(threadIdx.x & ~31u);
        const uint4 *rec4 = reinterpret_cast<const uint4 *>(sA + bytes_a);
        for (int r0 = wrow; r0 < nr; r0 += blockDim.x) {
            const int sb = (int)(rec4[r0].y & 0xffffu);
            const int se = (r0 + 32 < nr) ? (int)(rec4[r0 + 32].y & 0xffffu) : ns;
            __syncwarp();
            for (int s = sb + (int)(threadIdx.x & 31u); s < se; s += 32) SPFEM_FAN_ST(T.values + sG[s], sV[s]);
        }
    }
#else
    __syncthreads();
    /* every slot of the tile exactly once; a warp covers the CSR rows of its
       threads, so its stores fill whole 32-byte sectors                        */
    for (int s = threadIdx.x; s < ns; s += blockDim.x) SPFEM_FAN_ST(T.values + sG[s], sV[s]);
#endif
}
/* two instantiations: meshes whose vertices have at most 8 ring neighbours
   (fewer registers, higher occupancy) and the general case (<= 12)         */
extern "C" __global__ void __launch_bounds__(@FANTHREADS@, @FANMINB8@) @NAME@_fan8(const SpfemFanArgs T)
{
    @NAME@_fan_tile<8, false>(T);
}
/* ... and tiles of <= 256 vertices: 20-byte records (8-bit ids, 4-bit positions) */
extern "C" __global__ void __launch_bounds__(@FANTHREADS@, @FANMINB8@) @NAME@_fan8t(const SpfemFanArgs T)
{
    @NAME@_fan_tile<8, true>(T);
}
extern "C" __global__ void __launch_bounds__(@FANTHREADS@, @FANMINB@) @NAME@_fan(const SpfemFanArgs T)
{
    @NAME@_fan_tile<12, false>(T);
}
#else
extern "C" void @NAME@_fan_host(const SpfemFanArgs *T, int ntiles, double *sV, int *sG, int maxn)
{
    for (int t = 0; t < ntiles; ++t) {
        const int *d = T->desc + 8 * (size_t)t;
        const unsigned char *src = T->blob + (size_t)(unsigned)d[0] * 16;
        for (int row = 0; row < d[3]; ++row) {
            if (maxn > 100)          /* 100 + 8: the 20-byte record format */
                @NAME@_fan_row<8, true>(T->base, (const double2 *)src, src + d[1], d[3], row, sV, sG);
            else if (maxn <= 8)
                @NAME@_fan_row<8, false>(T->base, (const double2 *)src, src + d[1], d[3], row, sV, sG);
            else
                @NAME@_fan_row<12, false>(T->base, (const double2 *)src, src + d[1], d[3], row, sV, sG);
        }
        for (int s = 0; s < d[4]; ++s) T->values[sG[s]] = sV[s];
    }
}
#endif
"""


def lit(x):
    """Exact C literal of a double."""
    x = float(x)
    if x != x:
        return "(0.0/0.0)"
    if x in (float("inf"), float("-inf")):
        return "(1.0/0.0)" if x > 0 else "(-1.0/0.0)"
    if x == int(x) and abs(x) < 1e15:
        return "%d.0" % int(x) if x >= 0 else "(-%d.0)" % int(-x)
    h = x.hex()
    return h if x >= 0 else "(%s)" % h


_FUN1 = {"abs": "fabs", "sin": "sin", "cos": "cos", "tan": "tan", "exp": "exp",
         "log": "log", "sqrt": "sqrt", "arcsin": "asin", "arccos": "acos",
         "arctan": "atan", "sinh": "sinh", "cosh": "cosh", "tanh": "tanh",
         "log10": "log10", "log2": "log2", "floor": "floor", "ceil": "ceil"}
_CMP = {"lt": "<", "le": "<=", "gt": ">", "ge": ">=", "eq": "==", "ne": "!="}


MAX_PARAMS = 32


def structural_constant(v):
    """Constants that are part of a form's structure (0, +-1, +-2, 1/2, 1/4, ...
    -- what sym-grad / trace / identity expressions are made of) stay literals;
    every other captured Python scalar (a time step, a penalty, a material
    parameter, ``np.pi`` multiples) becomes a run-time parameter, so a time loop
    or a continuation over such a value re-uses one compiled kernel."""
    v = float(v)
    return v == v and abs(v) <= 4.0 and (v * 4.0) == int(v * 4.0)


class Emitter(object):
    """Emits the coefficient DAG as straight-line C; ``scope`` separates the
    copies of point-dependent nodes (one per quadrature point).  ``params``
    collects the values of the non-structural constants in the order of their
    first use: the generated code reads them from ``P.params[i]``."""

    def __init__(self, adjugate=False):
        self.lines = []
        self.done = {}
        self.params = []
        self._pidx = {}
        self._n = 0
        # affine interior kernels: |detA| * (sum of products of two entries of
        # A^{-1}) is emitted as (the same sum over the adjugate) / |detA| -- the
        # stiffness-type coefficients G = |detA| A^{-1} A^{-T} without forming A^{-1}
        self.adjugate = adjugate
        self._quad = {}
        self._done_adj = {}

    def _is_quad(self, node):
        """``node`` is homogeneous of degree two in the ``iA`` symbols."""
        r = self._quad.get(node.uid)
        if r is None:
            k, a = node.kind, node.args
            if k == "mul":
                iA = lambda x: x.kind == "sym" and x.value.startswith("iA") and not x.qdep
                cst = lambda x: x.kind == "const" and structural_constant(x.value)
                r = ((iA(a[0]) and iA(a[1])) or (cst(a[0]) and self._is_quad(a[1]))
                     or (cst(a[1]) and self._is_quad(a[0])))
            elif k in ("add", "sub"):
                r = self._is_quad(a[0]) and self._is_quad(a[1])
            elif k == "neg":
                r = self._is_quad(a[0])
            else:
                r = False
            self._quad[node.uid] = r
        return r

    def _name_adj(self, node, symname):
        """Emit a quadratic ``iA`` expression over the adjugate symbols ``jA``."""
        nm = self._done_adj.get(node.uid)
        if nm is not None:
            return nm
        if node.kind == "sym":
            return "jA_" + node.value[2:]
        if node.kind == "const":
            return lit(node.value)
        a = [self._name_adj(x, symname) for x in node.args]
        e = {"mul": "%s * %s", "add": "%s + %s", "sub": "%s - %s", "neg": "-%s"}[node.kind] % tuple(a)
        nm = "c%d" % self._n
        self._n += 1
        self.emit("const double %s = %s;" % (nm, e))
        self._done_adj[node.uid] = nm
        return nm

    def emit(self, s):
        self.lines.append("    " + s)

    def name(self, node, scope, symname):
        if node.kind == "const":
            if (structural_constant(node.value) or os.environ.get("SPFEM_PARAMS", "1") == "0"
                    or node.value != node.value or abs(node.value) == float("inf")):
                return lit(node.value)
            i = self._pidx.get(node.uid)
            if i is None:
                if len(self.params) >= MAX_PARAMS:
                    return lit(node.value)
                i = self._pidx[node.uid] = len(self.params)
                self.params.append(float(node.value))
            return "P.params[%d]" % i
        sc = scope if node.qdep else ""
        key = (node.uid, sc)
        nm = self.done.get(key)
        if nm is not None:
            return nm
        if node.kind == "sym":
            nm = symname(node.value, sc)
            self.done[key] = nm
            return nm
        k = node.kind
        if k == "mul" and self.adjugate:
            for i in (0, 1):
                sy, q = node.args[i], node.args[1 - i]
                if sy.kind == "sym" and sy.value == "absdet" and self._is_quad(q):
                    nm = "c%d" % self._n
                    self._n += 1
                    self.emit("const double %s = %s * rabsdet;" % (nm, self._name_adj(q, symname)))
                    self.done[key] = nm
                    return nm
        a = [self.name(x, scope, symname) for x in node.args]
        if k == "add":
            e = "%s + %s" % (a[0], a[1])
        elif k == "sub":
            e = "%s - %s" % (a[0], a[1])
        elif k == "mul":
            e = "%s * %s" % (a[0], a[1])
        elif k == "div":
            e = "%s / %s" % (a[0], a[1])
        elif k == "neg":
            e = "-%s" % a[0]
        elif k == "pow":
            e = "pow(%s, %s)" % (a[0], a[1])
        elif k in _FUN1:
            e = "%s(%s)" % (_FUN1[k], a[0])
        elif k in _CMP:
            e = "((%s %s %s) ? 1.0 : 0.0)" % (a[0], _CMP[k], a[1])
        elif k == "and":
            e = "((%s != 0.0 && %s != 0.0) ? 1.0 : 0.0)" % (a[0], a[1])
        elif k == "or":
            e = "((%s != 0.0 || %s != 0.0) ? 1.0 : 0.0)" % (a[0], a[1])
        elif k == "not":
            e = "((%s != 0.0) ? 0.0 : 1.0)" % a[0]
        elif k == "max":
            e = "fmax(%s, %s)" % (a[0], a[1])
        elif k == "min":
            e = "fmin(%s, %s)" % (a[0], a[1])
        elif k == "arctan2":
            e = "atan2(%s, %s)" % (a[0], a[1])
        elif k == "sign":
            e = "((%s > 0.0) ? 1.0 : ((%s < 0.0) ? -1.0 : 0.0))" % (a[0], a[0])
        else:
            raise NotImplementedError("codegen: operator %s" % k)
        # numbered in emission order (not by the interned node id): the same form
        # traced twice gives the same text, hence the same kernel-cache key
        nm = "c%d%s" % (self._n, ("_" + sc) if sc else "")
        self._n += 1
        self.emit("const double %s = %s;" % (nm, e))
        self.done[key] = nm
        return nm


def collect_syms(forms):
    """Names of all symbols used by the coefficient DAGs of ``forms``."""
    seen, syms = set(), set()
    stack = [c for f in forms for c in f.terms.values()]
    while stack:
        n = stack.pop()
        if n.uid in seen:
            continue
        seen.add(n.uid)
        if n.kind == "sym":
            syms.add(n.value)
        stack.extend(n.args)
    return syms


# ---------------------------------------------------------------------------
# geometry text
# ---------------------------------------------------------------------------
def _mul(a, b):
    return "__dmul_rn(%s, %s)" % (a, b)


def _add(a, b):
    return "__dadd_rn(%s, %s)" % (a, b)


def _sub(a, b):
    return "__dsub_rn(%s, %s)" % (a, b)


def _div(a, b):
    return "__ddiv_rn(%s, %s)" % (a, b)


def _neg(a):
    return "(-%s)" % a


def affine_geometry(dim, sfx="", elem="e", vt="P.t", ldv="P.ldt",
                    fast_inverse=False, pc="P.p", ldpc="P.ldp", tiled_pairs=False,
                    coords_given=False):
    """C statements computing ``A, b, detA, invA`` of element ``elem`` with
    the operation order of spfem/mapping.py:207-229 (tri) / 244-279 (tet).
    Variables carry the suffix ``sfx`` (facet kernels have two elements)."""
    L = []
    nve = dim + 1
    for k in range(nve):
        if coords_given:
            break
        L.append("const int vtx%s_%d = %s[%d * %s + %s];" % (sfx, k, vt, k, ldv, elem))
    # vertex coordinates X{sfx}_{i}_{k}: coordinate i of local vertex k
    for k in range(nve):
        if coords_given:
            break
        if tiled_pairs and dim == 2:
            L.append("double X%s_0_%d, X%s_1_%d;" % (sfx, k, sfx, k))
            L.append("if (TILED) { const double2 cxy = reinterpret_cast<const double2 *>(%s)[vtx%s_%d]; "
                     "X%s_0_%d = cxy.x; X%s_1_%d = cxy.y; } else { X%s_0_%d = %s[vtx%s_%d]; "
                     "X%s_1_%d = %s[%s + vtx%s_%d]; }"
                     % (pc, sfx, k, sfx, k, sfx, k, sfx, k, pc, sfx, k, sfx, k, pc, ldpc, sfx, k))
        else:
            for i in range(dim):
                L.append("const double X%s_%d_%d = %s[%d * %s + vtx%s_%d];"
                         % (sfx, i, k, pc, i, ldpc, sfx, k))
    for i in range(dim):
        L.append("const double b%s_%d = X%s_%d_0;" % (sfx, i, sfx, i))
        for j in range(dim):
            L.append("const double A%s_%d%d = %s;" % (
                sfx, i, j, _sub("X%s_%d_%d" % (sfx, i, j + 1), "b%s_%d" % (sfx, i))))
    A = lambda i, j: "A%s_%d%d" % (sfx, i, j)
    if dim == 2:
        L.append("const double detA%s = %s;" % (
            sfx, _sub(_mul(A(0, 0), A(1, 1)), _mul(A(0, 1), A(1, 0)))))
        inv = {(0, 0): A(1, 1), (0, 1): _neg(A(0, 1)),
               (1, 0): _neg(A(1, 0)), (1, 1): A(0, 0)}
    else:
        def m2(a, b, c, d):   # a*b - c*d
            return _sub(_mul(a, b), _mul(c, d))
        # detA = A00*(A11*A22 - A12*A21) - A01*(A10*A22 - A12*A20)
        #      + A02*(A10*A21 - A11*A20)          (spfem/mapping.py:265-267)
        t0 = _mul(A(0, 0), m2(A(1, 1), A(2, 2), A(1, 2), A(2, 1)))
        t1 = _mul(A(0, 1), m2(A(1, 0), A(2, 2), A(1, 2), A(2, 0)))
        t2 = _mul(A(0, 2), m2(A(1, 0), A(2, 1), A(1, 1), A(2, 0)))
        L.append("const double detA%s = %s;" % (sfx, _add(_sub(t0, t1), t2)))

        def cof(a, b, c, d):  # (-a*b + c*d)   (spfem/mapping.py:271-279)
            return _add(_mul(_neg(a), b), _mul(c, d))
        inv = {
            (0, 0): cof(A(1, 2), A(2, 1), A(1, 1), A(2, 2)),
            (1, 0): m2(A(1, 2), A(2, 0), A(1, 0), A(2, 2)),
            (2, 0): cof(A(1, 1), A(2, 0), A(1, 0), A(2, 1)),
            (0, 1): m2(A(0, 2), A(2, 1), A(0, 1), A(2, 2)),
            (1, 1): cof(A(0, 2), A(2, 0), A(0, 0), A(2, 2)),
            (2, 1): m2(A(0, 1), A(2, 0), A(0, 0), A(2, 1)),
            (0, 2): cof(A(0, 2), A(1, 1), A(0, 1), A(1, 2)),
            (1, 2): m2(A(0, 2), A(1, 0), A(0, 0), A(1, 2)),
            (2, 2): cof(A(0, 1), A(1, 0), A(0, 0), A(1, 1)),
        }
    if fast_inverse:
        # one division: adj * (1/det) instead of adj / det (<= 1 ulp apart)
        L.append("const double rdetA%s = spfem_rcp(detA%s);" % (sfx, sfx))
        L.append("const double rabsdet%s = fabs(rdetA%s);" % (sfx, sfx))
        for (i, j), e in sorted(inv.items()):
            L.append("const double jA%s_%d%d = %s;" % (sfx, i, j, e))
            L.append("const double iA%s_%d%d = jA%s_%d%d * rdetA%s;" % (sfx, i, j, sfx, i, j, sfx))
    else:
        for (i, j), e in sorted(inv.items()):
            L.append("const double iA%s_%d%d = %s;" % (sfx, i, j,
                                                      _div(e, "detA%s" % sfx)))
    L.append("const double absdet%s = fabs(detA%s);" % (sfx, sfx))
    return L


def q1_geometry_q(tag, xq, yq):
    """Jacobian of the bilinear map at the reference point ``(xq, yq)``
    (spfem/mapping.py:58-74, 142-171); the corner coordinates ``Q{i}_{k}``
    must be in scope.  ``xq``/``yq`` are floats (folded into literals);
    variables get the suffix ``tag``."""
    L = []
    one_m_y, one_p_y = lit(1.0 - yq), lit(1.0 + yq)
    one_m_x, one_p_x = lit(1.0 - xq), lit(1.0 + xq)
    q = tag
    for i in range(2):
        P = ["Q%d_%d" % (i, k) for k in range(4)]
        # 0.25*(-P0*(1-y) + P1*(1-y) + P2*(1+y) - P3*(1+y))
        e0 = _sub(_add(_add(_mul(_neg(P[0]), one_m_y), _mul(P[1], one_m_y)),
                       _mul(P[2], one_p_y)), _mul(P[3], one_p_y))
        # 0.25*(-P0*(1-x) - P1*(1+x) + P2*(1+x) + P3*(1-x))
        e1 = _add(_add(_sub(_mul(_neg(P[0]), one_m_x), _mul(P[1], one_p_x)),
                       _mul(P[2], one_p_x)), _mul(P[3], one_m_x))
        L.append("const double J%d0_%s = %s;" % (i, q, _mul("0.25", e0)))
        L.append("const double J%d1_%s = %s;" % (i, q, _mul("0.25", e1)))
    L.append("const double detJ_%s = %s;" % (
        q, _sub(_mul("J00_%s" % q, "J11_%s" % q), _mul("J01_%s" % q, "J10_%s" % q))))
    L.append("const double iJ_00_%s = %s;" % (q, _div("J11_%s" % q, "detJ_%s" % q)))
    L.append("const double iJ_01_%s = %s;" % (q, _div(_neg("J01_%s" % q), "detJ_%s" % q)))
    L.append("const double iJ_10_%s = %s;" % (q, _div(_neg("J10_%s" % q), "detJ_%s" % q)))
    L.append("const double iJ_11_%s = %s;" % (q, _div("J00_%s" % q, "detJ_%s" % q)))
    L.append("const double absdetq_%s = fabs(detJ_%s);" % (q, q))
    return L


def _q1_basis(x, y, k):
    sx = (-1, 1, 1, -1)[k]
    sy = (-1, -1, 1, 1)[k]
    return 0.25 * (1 + sx * x) * (1 + sy * y)


# ---------------------------------------------------------------------------
# interior kernels
# ---------------------------------------------------------------------------
class InteriorSpec(object):
    """Everything the generator needs for one ``iasm`` kernel.

    ``blocks``: ``{(cu, cv): Form}``; ``tab_u`` / ``tab_v``: reference tables
    ``psi[a][j][q]`` with ``a = 0`` the value and ``1 + l`` the l-th reference
    derivative of the *scalar* parent element."""

    def __init__(self, refdom, dim, bilinear, nbu, nbv, ncu, ncv, blocks,
                 X, W, tab_u, tab_v, nw=0, nfld=0):
        self.refdom, self.dim, self.bilinear = refdom, dim, bilinear
        self.nbu, self.nbv, self.ncu, self.ncv = nbu, nbv, ncu, ncv
        self.blocks, self.X, self.W = blocks, X, W
        self.tab_u, self.tab_v, self.nw = tab_u, tab_v, nw
        self.nfld = nfld            # host-evaluated coefficient fields (tracer.FieldContext)

    @property
    def needs_el(self):
        """The kernel indexes per-element data (``interp`` gathers, fields)."""
        return self.nw > 0 or self.nfld > 0

    @property
    def n_rows(self):
        return self.nbv * self.ncv

    @property
    def n_cols(self):
        return self.nbu * self.ncu if self.bilinear else 1

    @property
    def n_entries(self):
        return self.n_rows * self.n_cols


def _psi(tab, a, j, q):
    return 1.0 if a == NONE else tab[a][j][q]


def interior_terms(spec, drop_tol=1e-15):
    """For every local entry the list ``[(constant, coef_node, scope)]``.

    Returns ``(entries, coef_uses)`` where ``entries[idx]`` follows the
    reference's COO block numbering ``idx = Nbfun_v*j + i``
    (spfem/assembly.py:977) and ``coef_uses`` is the ordered list of
    ``(node, scope)`` that must be evaluated."""
    nq = len(spec.W)
    entries = [[] for _ in range(spec.n_entries)]
    uses, seen = [], set()
    nbu = spec.nbu if spec.bilinear else 1
    for (cu, cv), form in sorted(spec.blocks.items()):
        for (a, b), coef in sorted(form.terms.items(),
                                   key=lambda kv: kv[0]):
            if a != NONE and not spec.bilinear:
                raise ValueError("trial function in a linear form")
            # K[j, i, q] = W_q psi^a_j(q) psi^b_i(q)
            K = np.empty((nbu, spec.nbv, nq))
            for j in range(nbu):
                for i in range(spec.nbv):
                    for q in range(nq):
                        # the product of the two basis values first: commutative, so
                        # the transposed entry of a symmetric form gets bit-identical
                        # constants and is recognised as an alias below
                        K[j, i, q] = spec.W[q] * (_psi(spec.tab_u, a, j, q)
                                                  * _psi(spec.tab_v, b, i, q))
            if coef.qdep:
                scale = np.max(np.abs(K)) if K.size else 0.0
                for q in range(nq):
                    used = False
                    for j in range(nbu):
                        for i in range(spec.nbv):
                            if abs(K[j, i, q]) > drop_tol * scale:
                                idx = ((j * spec.ncu + cu) * spec.n_rows
                                       + (i * spec.ncv + cv))
                                entries[idx].append((K[j, i, q], coef, "%d" % q))
                                used = True
                    if used and (coef.uid, q) not in seen:
                        seen.add((coef.uid, q))
                        uses.append((coef, "%d" % q))
            else:
                # exactly rounded sum over the points (math.fsum)
                import math
                Ks = np.array([[math.fsum(K[j, i, :]) for i in range(spec.nbv)]
                               for j in range(nbu)])
                mag = np.array([[math.fsum(np.abs(K[j, i, :]))
                                 for i in range(spec.nbv)] for j in range(nbu)])
                scale = np.max(mag) if mag.size else 0.0
                used = False
                for j in range(nbu):
                    for i in range(spec.nbv):
                        if abs(Ks[j, i]) > drop_tol * scale:
                            idx = ((j * spec.ncu + cu) * spec.n_rows
                                   + (i * spec.ncv + cv))
                            entries[idx].append((Ks[j, i], coef, ""))
                            used = True
                if used and (coef.uid, "") not in seen:
                    seen.add((coef.uid, ""))
                    uses.append((coef, ""))
    return entries, uses


def _ctable(name, arr):
    """A constant table: ``__constant__`` memory on the device (uniform reads:
    every thread of a warp is at the same quadrature point), a static array in
    the host emulation."""
    arr = np.asarray(arr, dtype=np.float64)
    dims = "".join("[%d]" % n for n in arr.shape)

    def braces(a):
        if a.ndim == 1:
            return "{" + ", ".join(lit(v) for v in a) + "}"
        return "{" + ", ".join(braces(x) for x in a) + "}"
    return "SPFEM_CONST double %s%s = %s;" % (name, dims, braces(arr))


def looped_quad_body(spec, syms, em):
    """Element body for ``MappingQ1`` meshes as a LOOP over the quadrature points
    (spfem/mapping.py:47-191, spfem/assembly.py:961-984).

    The straight-line generator folds ``W_q psi_j(X_q) psi_i(X_q)`` into one
    constant per (entry, point, coefficient): 45 x 16 x 3 terms for Q2 Laplace,
    about 100 KB of code that no instruction cache holds.  Here the reference
    tables live in constant memory; per point the bilinear Jacobian and the
    coefficient DAG are evaluated once, then

        H^b_j = sum_a G^{ab} (W_q psi^a_j)        (per trial function and b)
        A[j, i] += sum_b H^b_j psi^b_i            (distinct entries only)

    Elements that are parallelograms take a loop-free branch (constant Jacobian:
    reference integrals folded into literals) when the form's point data is
    geometry only; ``SPFEM_PARALLELOGRAMS`` compiles the body without the loop
    for meshes that consist of parallelograms only (``spec.affine_variant``).

    Returns ``(tables, body, alias, n_unique, n_terms)``."""
    dim, nq = spec.dim, len(spec.W)
    nbu = spec.nbu if spec.bilinear else 1
    X = spec.X
    tables = [_ctable("SPFEM_QT", [1.0 - X[1], 1.0 + X[1], 1.0 - X[0], 1.0 + X[0]]),
              _ctable("SPFEM_QN", [[_q1_basis(X[0, q], X[1, q], c) for q in range(nq)]
                                   for c in range(4)]),
              _ctable("SPFEM_QW", spec.W),
              _ctable("SPFEM_TV", np.asarray(spec.tab_v, dtype=np.float64))]
    if spec.bilinear:
        tables.append(_ctable("SPFEM_TUW", np.asarray(spec.tab_u, dtype=np.float64)
                              * np.asarray(spec.W)[None, None, :]))
    if spec.nw:
        tables.append(_ctable("SPFEM_TU0", np.asarray(spec.tab_u[0], dtype=np.float64)))
    body = []
    for k in range(4):
        body.append("const int vtx_%d = vt[%d * ldv + ev];" % (k, k))
    for k in range(4):
        body.append("double Q0_%d, Q1_%d;" % (k, k))
        body.append("if (TILED) { const double2 cxy = reinterpret_cast<const double2 *>(pc)[vtx_%d]; "
                    "Q0_%d = cxy.x; Q1_%d = cxy.y; } else { Q0_%d = pc[vtx_%d]; "
                    "Q1_%d = pc[ldpc + vtx_%d]; }" % (k, k, k, k, k, k, k))
    for sw in range(spec.nw):
        if ("w%d" % sw) in syms:
            for j in range(spec.nbu):
                body.append("const double wc%d_%d = P.interp[%d][P.tdof_u[%d * P.ldd + e]];"
                            % (sw, j, sw, j))
    if any(nm.startswith("dw") for nm in syms):
        raise NotImplementedError("gradients of interpolated fields on quadrilaterals")

    # ---- entries -> term lists (a, b, coefficient), distinct entries -------------
    same_tab = spec.bilinear and spec.tab_u is spec.tab_v or (
        spec.bilinear and np.array_equal(np.asarray(spec.tab_u), np.asarray(spec.tab_v)))
    tu = np.asarray(spec.tab_u, dtype=np.float64) if spec.bilinear else None
    tv = np.asarray(spec.tab_v, dtype=np.float64)
    eterms = [[] for _ in range(spec.n_entries)]
    for (cu, cv), form in sorted(spec.blocks.items()):
        for (a, b), coef in sorted(form.terms.items(), key=lambda kv: kv[0]):
            if a != NONE and not spec.bilinear:
                raise ValueError("trial function in a linear form")
            for j in range(nbu):
                if a != NONE and not np.any(tu[a][j]):
                    continue
                for i in range(spec.nbv):
                    if b != NONE and not np.any(tv[b][i]):
                        continue
                    idx = (j * spec.ncu + cu) * spec.n_rows + (i * spec.ncv + cv)
                    eterms[idx].append((a, j, b, i, coef, cu, cv))
    groups, order_u, alias = {}, [], [0] * spec.n_entries
    for idx, terms in enumerate(eterms):
        key = []
        for a, j, b, i, coef, cu, cv in terms:
            f1, f2 = ("u", a, j), ("v", b, i)
            if same_tab:                   # psi^a_j psi^b_i is symmetric in the two factors
                f1, f2 = sorted([(a, j), (b, i)])
            key.append((coef.uid, f1, f2))
        gkey = tuple(sorted(key))
        if gkey not in groups:
            groups[gkey] = (len(order_u), idx, [])
            order_u.append(gkey)
        groups[gkey][2].append(idx)
        alias[idx] = groups[gkey][0]
    n_unique = len(order_u)
    npart = int(os.environ.get("SPFEM_TILE_NPART", "0"))
    if npart <= 0:
        npart = 1 if n_unique <= 24 else 2
    spec.npart = npart
    # The distinct entries are dealt to the NPART threads of an element BY TRIAL
    # FUNCTION (greedy balance): a thread then needs the H vectors of its own trial
    # functions only, not all of them.
    part_of_u = {}
    if npart > 1:
        by_j = {}
        for gkey in order_u:
            u, idx, _ = groups[gkey]
            jj = eterms[idx][0][1] if eterms[idx] else 0
            by_j.setdefault(jj, []).append(u)
        load = [0] * npart
        for jj in sorted(by_j, key=lambda k: -len(by_j[k])):
            tgt = load.index(min(load))
            load[tgt] += len(by_j[jj])
            for u in by_j[jj]:
                part_of_u[u] = tgt
    for gkey in order_u:
        u = groups[gkey][0]
        part_of_u.setdefault(u, 0 if npart == 1 else u % npart)
        body.append("double acc_%d = 0.0;" % u)

    # ---- the loop over the quadrature points --------------------------------------
    # Parallelograms (every element of a uniform grid) have a constant Jacobian: if the
    # form's point data is geometry only (no x, w, fields), such an element takes a
    # loop-free branch with the reference integrals folded into literals (below).
    # The test is exact (P0 - P1 + P2 - P3 == 0); anything else takes the loop.
    geom_only = not any(nm[0] in "xwf" for nm in syms if nm not in ("hq",))
    pre = []
    if geom_only and os.environ.get("SPFEM_QAFFINE", "1") != "0":
        # (SPFEM_PARALLELOGRAMS: the engine found every element of the mesh to be one --
        # the kernel variant without the loop branch needs half the registers)
        pre.append("#ifdef SPFEM_PARALLELOGRAMS")
        pre.append("const bool spfem_affine = true;")
        pre.append("#else")
        pre.append("const bool spfem_affine = (Q0_0 - Q0_1 + Q0_2 - Q0_3 == 0.0) && "
                   "(Q1_0 - Q1_1 + Q1_2 - Q1_3 == 0.0);")
        pre.append("#endif")
    L = ["for (int q = 0; q < %d; ++q) {" % nq]
    G = []          # per-point geometry + coefficients
    L_ = L
    L = G
    L.append("const double omy = SPFEM_QT[0][q], opy = SPFEM_QT[1][q], omx = SPFEM_QT[2][q], "
             "opx = SPFEM_QT[3][q];")
    # (plain products / sums: the compiler may contract them into FMAs -- unlike the
    # coordinates x, which masks compare with ==, the Jacobian needs no bit-exactness)
    for i in range(2):
        P = ["Q%d_%d" % (i, k) for k in range(4)]
        L.append("const double J%d0_L = 0.25 * ((%s - %s) * omy + (%s - %s) * opy);"
                 % (i, P[1], P[0], P[2], P[3]))
        L.append("const double J%d1_L = 0.25 * ((%s - %s) * omx + (%s - %s) * opx);"
                 % (i, P[3], P[0], P[2], P[1]))
    L.append("const double detJ_L = J00_L * J11_L - J01_L * J10_L;")
    # one division per point: adj * (1 / det) instead of adj / det (<= 1 ulp apart,
    # like the affine path; on axis-aligned meshes J01 = J10 = 0 exactly and a
    # zero numerator would send every one of those divisions down the slow path)
    L.append("const double rdetJ_L = 1.0 / detJ_L;")
    L.append("const double iJ_00_L = J11_L * rdetJ_L;")
    L.append("const double iJ_01_L = (-J01_L) * rdetJ_L;")
    L.append("const double iJ_10_L = (-J10_L) * rdetJ_L;")
    L.append("const double iJ_11_L = J00_L * rdetJ_L;")
    L.append("const double absdetq_L = fabs(detJ_L);")
    if "hq" in syms:
        L.append("const double hq_L = pow(absdetq_L, %s);" % lit(1.0 / dim))
    for k in range(dim):
        if ("x%d" % k) in syms:
            acc = _mul("Q%d_0" % k, "SPFEM_QN[0][q]")
            for c in range(1, 4):
                acc = _add(acc, _mul("Q%d_%d" % (k, c), "SPFEM_QN[%d][q]" % c))
            L.append("const double x%d_L = %s;" % (k, acc))
    for sw in range(spec.nw):
        if ("w%d" % sw) in syms:
            L.append("const double w%d_L = 0.0 + %s;" % (sw, " + ".join(
                "wc%d_%d * SPFEM_TU0[%d][q]" % (sw, j, j) for j in range(spec.nbu))))
    for i in range(spec.nfld):
        if ("fld%d" % i) in syms:
            L.append("const double fld%d_L = P.fields[%d][(long long)q * P.ldf + e];" % (i, i))

    def symname(nm, scope):
        if nm.startswith("iJ"):
            return "iJ_%s_L" % nm[2:]
        return "%s_L" % nm if scope else nm

    # coefficient DAG, once (element-constant sub-expressions are hoisted by the compiler)
    cname = {}
    for terms in eterms:
        for a, j, b, i, coef, cu, cv in terms:
            if coef.uid not in cname:
                cname[coef.uid] = em.name(coef, "L", symname)
    L += [ln.strip() for ln in em.lines]
    L = L_
    L += G
    affine = []
    if pre:
        # Parallelogram: the coefficients are constants of the element, so an entry is
        #     sum_terms c * (sum_q W_q psi^a_j(X_q) psi^b_i(X_q))
        # with the reference integrals folded into literals (math.fsum), exactly what
        # the straight-line generator does for affine simplices: no loop, no tables.
        import math
        affine.append("if (spfem_affine) {")
        affine.append("const int q = 0;")
        affine += G
        folded, fmag = {}, 0.0
        for gkey in order_u:
            u, idx, _ = groups[gkey]
            merged, order = {}, []
            for a, j, b, i, coef, cu, cv in eterms[idx]:
                prod = [spec.W[q] * ((1.0 if a == NONE else tu[a][j][q])
                                     * (1.0 if b == NONE else tv[b][i][q])) for q in range(nq)]
                fmag = max(fmag, math.fsum(abs(x) for x in prod))
                if coef.uid not in merged:
                    merged[coef.uid] = 0.0
                    order.append(coef.uid)
                merged[coef.uid] = math.fsum([merged[coef.uid]] + prod)
            folded[u] = [(merged[c], c) for c in order]
        for gkey in order_u:
            u = groups[gkey][0]
            terms = ["%s * %s" % (lit(K), cname[c]) for K, c in folded[u] if abs(K) > 1e-15 * fmag]
            if terms:
                affine.append("if (PART < 0 || PART == %d) acc_%d = %s;"
                              % (part_of_u[u], u, " + ".join(terms)))
        affine.append("}")
        affine.append("#ifndef SPFEM_PARALLELOGRAMS")
        affine.append("else")
    # H^b_j per (block, trial function, b) that some distinct entry needs
    reps = [groups[g][1] for g in order_u]
    need_h, hterms = [], {}
    for idx in reps:
        for a, j, b, i, coef, cu, cv in eterms[idx]:
            hk = (cu, cv, j, b)
            if hk not in hterms:
                hterms[hk] = []
                need_h.append(hk)
    for (cu, cv), form in sorted(spec.blocks.items()):
        for (a, b), coef in sorted(form.terms.items(), key=lambda kv: kv[0]):
            for j in range(nbu):
                hk = (cu, cv, j, b)
                if hk in hterms and not (a != NONE and not np.any(tu[a][j])):
                    fac = "SPFEM_QW[q]" if a == NONE else "SPFEM_TUW[%d][%d][q]" % (a, j)
                    hterms[hk].append("%s * %s" % (cname[coef.uid], fac))
    n_terms = 0
    hname = {}
    # which parts need which H vector
    h_parts = {}
    for gkey in order_u:
        u, idx, idxs = groups[gkey]
        for a, j, b, i, coef, cu, cv in eterms[idx]:
            h_parts.setdefault((cu, cv, j, b), set()).add(part_of_u[u])
    for hk in need_h:
        cu, cv, j, b = hk
        hname[hk] = "H_%d_%d_%d_%s" % (cu, cv, j, "n" if b == NONE else str(b))
        cond = " || ".join("PART == %d" % pp for pp in sorted(h_parts.get(hk, {0})))
        L.append("double %s = 0.0; if (PART < 0 || %s) %s = %s;"
                 % (hname[hk], cond, hname[hk], " + ".join(hterms[hk]) or "0.0"))
        n_terms += len(hterms[hk])
    for gkey in order_u:
        u, idx, idxs = groups[gkey]
        by_b = {}
        for a, j, b, i, coef, cu, cv in eterms[idx]:
            by_b.setdefault((cu, cv, j, b), i)
        stmts = []
        for (cu, cv, j, b), i in by_b.items():
            # one statement per product: each contracts into a single FMA
            stmts.append("acc_%d += %s;" % (u, hname[(cu, cv, j, b)] if b == NONE
                         else "%s * SPFEM_TV[%d][%d][q]" % (hname[(cu, cv, j, b)], b, i)))
        n_terms += len(stmts)
        if stmts:
            L.append("if (PART < 0 || PART == %d) { %s }" % (part_of_u[u], " ".join(stmts)))
    L.append("}")
    if affine:
        L.append("#endif")
    spec.affine_variant = bool(affine)
    body += pre + affine + L
    for gkey in order_u:
        u, idx, idxs = groups[gkey]
        body.append("if (PART < 0 || PART == %d) { const double val = acc_%d;" % (part_of_u[u], u))
        body.append("  if (TILED) { out[(long long)%d * ldo + k] = val; } else {" % u)
        for ii in idxs:
            body.append("    SPFEM_PUT(%d, val);" % ii)
        body.append("  } }")
    return tables, body, alias, n_unique, n_terms * nq // 4 + 40


def generate_interior(spec, name="spfem_elem"):
    """Source text of the interior element kernel for ``spec``."""
    dim, nq = spec.dim, len(spec.W)
    if spec.refdom == "quad" and os.environ.get("SPFEM_QLOOP", "1") != "0":
        syms = collect_syms(list(spec.blocks.values()))
        if not any(nm.startswith("dw") for nm in syms):   # (inorm's gradients: straight-line path)
            return _generate_interior_source(spec, name, looped=True)
    return _generate_interior_source(spec, name, looped=False)


def _generate_interior_source(spec, name, looped):
    dim, nq = spec.dim, len(spec.W)
    entries, uses = interior_terms(spec)
    syms = collect_syms(list(spec.blocks.values()))
    quad = spec.refdom == "quad"
    em = Emitter(adjugate=not quad and os.environ.get("SPFEM_ADJUGATE", "1") != "0")

    body = []
    tables = []
    if looped:
        tables, body, l_alias, l_nuniq, l_nterms = looped_quad_body(spec, syms, em)
    if looped:
        pass
    elif quad:
        for k in range(4):
            body.append("const int vtx_%d = vt[%d * ldv + ev];" % (k, k))
        for k in range(4):
            body.append("double Q0_%d, Q1_%d;" % (k, k))
            body.append("if (TILED) { const double2 cxy = reinterpret_cast<const double2 *>(pc)[vtx_%d]; "
                        "Q0_%d = cxy.x; Q1_%d = cxy.y; } else { Q0_%d = pc[vtx_%d]; "
                        "Q1_%d = pc[ldpc + vtx_%d]; }" % (k, k, k, k, k, k, k))
        for q in range(nq):
            body += q1_geometry_q("%d" % q, spec.X[0, q], spec.X[1, q])
            if "hq" in syms:
                body.append("const double hq_%d = pow(absdetq_%d, %s);"
                            % (q, q, lit(1.0 / dim)))
    else:
        body += affine_geometry(dim, vt="vt", ldv="ldv", elem="ev", fast_inverse=True,
                                pc="pc", ldpc="ldpc", tiled_pairs=True)
        n_geom_lines = len(body)
        if "h" in syms:
            body.append("const double h = pow(absdet, %s);" % lit(1.0 / dim))
    if looped:
        spec.alias, spec.n_unique, spec.n_terms = l_alias, l_nuniq, l_nterms
        alias = l_alias
        spec.params = list(em.params)
        spec.fan_ok = False
        core = []
    else:
        # global coordinates of the quadrature points (only if the form uses x)
        for k in range(dim):
            if ("x%d" % k) not in syms:
                continue
            for q in range(nq):
                if quad:
                    # sum_k P_k * N_k(X_q)                (spfem/mapping.py:104-110)
                    acc = _mul("Q%d_0" % k, lit(_q1_basis(spec.X[0, q], spec.X[1, q], 0)))
                    for c in range(1, 4):
                        acc = _add(acc, _mul("Q%d_%d" % (k, c), lit(
                            _q1_basis(spec.X[0, q], spec.X[1, q], c))))
                    body.append("const double x%d_%d = %s;" % (k, q, acc))
                else:
                    # outer(A_k0, X0) + outer(A_k1, X1) (+ ...) + b_k
                    #                                    (spfem/mapping.py:318-319)
                    acc = _mul("A_%d0" % k, lit(spec.X[0, q]))
                    for l in range(1, dim):
                        acc = _add(acc, _mul("A_%d%d" % (k, l), lit(spec.X[l, q])))
                    body.append("const double x%d_%d = %s;" % (k, q, _add(acc, "b_%d" % k)))
        # interpolated fields w[k] = sum_j interp[k][t_dof_u[j, e]] phi_j(X_q)
        #                                          (spfem/assembly.py:955-959)
        for s in range(spec.nw):
            use_w = ("w%d" % s) in syms
            use_dw = any(("dw%d_%d" % (s, a)) in syms for a in range(dim))
            if not (use_w or use_dw):
                continue
            for j in range(spec.nbu):
                body.append("const double wc%d_%d = P.interp[%d][P.tdof_u[%d * P.ldd + e]];"
                            % (s, j, s, j))
            for q in range(nq):
                if use_w:
                    terms = ["wc%d_%d * %s" % (s, j, lit(spec.tab_u[0][j][q]))
                             for j in range(spec.nbu)]
                    body.append("const double w%d_%d = 0.0 + %s;" % (s, q, " + ".join(terms)))
                if use_dw:
                    # gradient of the interpolated field: DF^{-T} applied to the
                    # reference gradient sum_j c_j dphi_j(X_q)  (inorm, assembly.py:1340-1346)
                    for l in range(dim):
                        terms = ["wc%d_%d * %s" % (s, j, lit(spec.tab_u[1 + l][j][q]))
                                 for j in range(spec.nbu)]
                        body.append("const double gw%d_%d_%d = 0.0 + %s;" % (s, l, q, " + ".join(terms)))
                    for a in range(dim):
                        inv = (lambda l: "iJ_%d%d_%d" % (l, a, q)) if quad else (lambda l: "iA_%d%d" % (l, a))
                        body.append("const double dw%d_%d_%d = %s;" % (
                            s, a, q, " + ".join("%s * gw%d_%d_%d" % (inv(l), s, l, q) for l in range(dim))))

        # host-evaluated coefficient fields: one value per element and quadrature point
        for i in range(spec.nfld):
            if ("fld%d" % i) in syms:
                for q in range(nq):
                    body.append("const double fld%d_%d = P.fields[%d][%dLL * P.ldf + e];" % (i, q, i, q))

        def symname(nm, scope):
            if nm.startswith("iA"):
                return "iA_" + nm[2:]
            if nm.startswith("iJ"):
                return "iJ_%s_%s" % (nm[2:], scope)
            if scope:
                return "%s_%s" % (nm, scope)
            return nm

        names = {}
        for node, scope in uses:
            names[(node.uid, scope)] = em.name(node, scope, symname)
        body += [ln.strip() for ln in em.lines]

        # Entries with identical term lists (symmetric forms: A_loc[j,i] == A_loc[i,j];
        # P1 mass: all off-diagonal entries) are *aliases*: the fused tile kernel
        # computes and stores each distinct value once (``alias[idx]`` = its slot in
        # shared memory); the two-kernel route stores every entry.
        groups, order_u, alias = {}, [], [0] * len(entries)
        for idx, terms in enumerate(entries):
            # merge terms that multiply the same coefficient (e.g. G^{12} == G^{21})
            merged, order = {}, []
            for K, node, scope in terms:
                key = (node.uid, scope)
                if key in merged:
                    merged[key] = (merged[key][0] + K, node, scope)
                else:
                    merged[key] = (K, node, scope)
                    order.append(key)
            terms = [merged[k] for k in order if merged[k][0] != 0.0]
            gkey = tuple(sorted((node.uid, scope, float(K).hex()) for K, node, scope in terms))
            if gkey not in groups:
                if terms:
                    expr = " + ".join("%s * %s" % (lit(K), names[(node.uid, scope)])
                                      for K, node, scope in terms)
                else:
                    expr = "0.0"
                groups[gkey] = (len(order_u), expr, [])
                order_u.append(gkey)
            groups[gkey][2].append(idx)
            alias[idx] = groups[gkey][0]
        # threads per element in the tile kernel's phase A (distinct values dealt
        # round-robin to the parts)
        # Measured (profiles/r02_owner_rule_sweep.jsonl): one thread per element for
        # forms with aliased entries (vector-P2 elasticity 63 of 144 values: 0.458 ->
        # 0.481, TetP2 49 of 100: 0.169 -> 0.183 with one part), two threads for large
        # local matrices without aliases (Stokes B, 36 of 36: 0.339 -> 0.364 with two)
        npart = int(os.environ.get("SPFEM_TILE_NPART", "0"))
        if npart <= 0:
            npart = 2 if (len(order_u) > 24 and len(order_u) == len(entries)) else 1
        spec.npart = npart
        for gkey in order_u:
            u, expr, idxs = groups[gkey]
            body.append("if (PART < 0 || PART == %d) { const double val = %s;" % (u % npart, expr))
            body.append("  if (TILED) { out[(long long)%d * ldo + k] = val; } else {" % u)
            for idx in idxs:
                body.append("    SPFEM_PUT(%d, val);" % idx)
            body.append("  } }")
        spec.alias = alias
        spec.n_unique = len(order_u)
        spec.params = list(em.params)
        # multiply-adds per element (cost estimate: is recomputing halo elements cheap?)
        spec.n_terms = sum(groups[g][1].count(" * ") for g in order_u) + len(em.lines)
        # Vertex-fan variant (TriP1-like: one DOF per vertex, element-constant
        # coefficients): the same computation with the vertex coordinates passed in
        spec.fan_ok = (spec.refdom == "tri" and spec.bilinear and spec.ncu == 1 and spec.ncv == 1
                       and spec.nbu == 3 and spec.nbv == 3
                       and not any(nm[0] in "xwdf" for nm in syms))
        core = []
        if spec.fan_ok:
            core = (affine_geometry(dim, fast_inverse=True, coords_given=True)
                    + body[n_geom_lines:])

    spec.looped = bool(looped)
    spec.minb = 3 if dim == 2 else 2          # CTAs of the tile kernel per SM (register cap)
    if looped and spec.n_unique > 24:
        spec.minb = 2
    # CTAs per SM of the variant for meshes of parallelograms only (no loop branch)
    spec.minb_affine = 3 if spec.n_unique > 24 else 4
    spec.affine_variant = bool(looped and getattr(spec, "affine_variant", False))
    src = [PRELUDE, DEVICE_HELPERS]
    if os.environ.get("SPFEM_RCP", "0") == "0":      # measured: the Newton reciprocal is not faster
        src.insert(0, "#define SPFEM_IEEE_RCP 1")
    if os.environ.get("SPFEM_FAN_NOCOMPUTE", "0") == "1":
        src.append("#define SPFEM_FAN_NOCOMPUTE 1")
    for flag in os.environ.get("SPFEM_FAN_X", "").split(","):    # kernel experiments (tools/fan_variants.py)
        if flag:
            src.append("#define SPFEM_FAN_%s 1" % flag)
    # small tables (Q1: 4 points) sit in constant memory; larger ones (Q2: 7.6 KB) would
    # thrash the constant cache -- global memory through L1 instead (measured, Q2: 2 x)
    tbytes = 8 * sum(t.count(",") + 1 for t in tables)
    space = os.environ.get("SPFEM_TABLE_SPACE", "constant" if tbytes <= 2048 else "global")
    space = "__constant__" if space == "constant" else "__device__ const"
    src.append("#ifdef SPFEM_HOST\n#define SPFEM_CONST static const\n#else\n"
               "#define SPFEM_CONST %s\n#endif" % space)
    src += tables
    src.append("#define SPFEM_NENTRIES %d" % spec.n_entries)
    src.append("#ifdef SPFEM_PARALLELOGRAMS\n#define SPFEM_TILE_MINB %s\n#else\n#define SPFEM_TILE_MINB %s\n#endif"
               % (os.environ.get("SPFEM_TILE_MINB_AFFINE", str(spec.minb_affine)),
                  os.environ.get("SPFEM_TILE_MINB", str(spec.minb))))
    src.append("#define SPFEM_NUNIQUE %d" % spec.n_unique)
    src.append("#define SPFEM_NPART %d" % spec.npart)
    src.append("#define SPFEM_NMASK %d" % ((spec.n_unique + 31) // 32))
    # block structure of a vector element: bsv x bsu values per scalar entry
    bsv, bsu = spec.ncv, (spec.ncu if spec.bilinear else 1)
    if os.environ.get("SPFEM_TILE_BLOCKS", "1") == "0":
        bsv = bsu = 1              # plan every CSR value as its own slot
    nbv_s, nbu_s = spec.n_rows // bsv, spec.n_cols // bsu
    ut = []
    for j in range(nbu_s):
        for i in range(nbv_s):
            for a_ in range(bsv):
                for b_ in range(bsu):
                    ut.append(alias[(bsu * j + b_) * spec.n_rows + (bsv * i + a_)])
    spec.bsv, spec.bsu, spec.ut = bsv, bsu, ut
    src.append("#define SPFEM_BSV %d" % bsv)
    src.append("#define SPFEM_BSU %d" % bsu)
    src.append("#define SPFEM_NB %d" % (bsv * bsu))
    src.append("#define SPFEM_NSIDX %d" % (nbv_s * nbu_s))
    # store of one local entry in the two-kernel route: entry-major planes
    src.append("#define SPFEM_PUT(idx, val) out[(long long)(idx) * ldo + k] = (val)")
    # body<TILED>(P, e, vt, ldv, ev, pc, ldpc, out, ldo, k): local matrix of
    # element e whose vertex ids are vt[r*ldv + ev]  ->  out[idx*ldo + k]
    # (TILED: only the distinct values, out[alias*ldo + k])
    # (PART >= 0: only the distinct values of that part; PART < 0: all)
    src.append("template <bool TILED, int PART, typename VT, typename OUT>")
    src.append("SPFEM_DEVICE void %s_body(const SpfemArgs &P, const long long e, "
               "const VT *__restrict__ vt, const long long ldv, const long long ev, "
               "const double *__restrict__ pc, const long long ldpc, "
               "OUT out, const long long ldo, const long long k)\n{" % name)
    src += ["    " + ln for ln in body]
    src.append("}")
    if spec.fan_ok:
        # core(P, x0,y0, x1,y1, x2,y2, out): distinct local values out[alias]
        src.append("SPFEM_DEVICE void %s_core(const SpfemArgs &P, const double X_0_0, "
                   "const double X_1_0, const double X_0_1, const double X_1_1, "
                   "const double X_0_2, const double X_1_2, double *__restrict__ out)\n{" % name)
        src.append("    const bool TILED = true; const int PART = -1; const long long ldo = 1, k = 0;")
        src += ["    " + ln for ln in core]
        src.append("}")
        # rows per tile = threads per CTA.  Forms that are cheap per triangle (mass: |detA|
        # times constants) are bound by the per-tile fixed cost (descriptor, TMA wait,
        # barrier): 160 rows per tile; forms with a real local row (Laplace: adjugate,
        # reciprocal, 62 registers) run best with 128 (measured, interleaved:
        # profiles/r02_fan_diagnostics.jsonl -- mass 0.1043 -> 0.1007 ms, Laplace
        # 0.0991 -> 0.102 ms with 160)
        heavy = spec.n_terms > 12
        spec.fan_rows = int(os.environ.get("SPFEM_FAN_ROWS", "128" if heavy else "160"))
        spec.fan_threads = int(os.environ.get("SPFEM_FAN_T", str(spec.fan_rows))) // 32 * 32
        minb8 = os.environ.get("SPFEM_FAN_MINB8", "7" if spec.fan_threads <= 128 else "5")
        src.append(FAN_KERNEL.replace("@NAME@", name).replace("@FANTHREADS@", str(spec.fan_threads))
                   .replace("@FANMINB8@", minb8)
                   .replace("@FANMINB@", os.environ.get("SPFEM_FAN_MINB", "6" if spec.fan_threads <= 128 else "4"))
                   .replace("@A00@", str(alias[0])).replace("@A10@", str(alias[3]))
                   .replace("@A20@", str(alias[6])))
    src.append("#ifndef SPFEM_HOST")
    src.append('extern "C" __global__ void __launch_bounds__(128) %s(const SpfemArgs P)\n{' % name)
    src.append("    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;")
    src.append("    if (k >= P.n) return;")
    src.append("    const long long e = P.sel ? (long long)P.sel[k] : k;")
    src.append("    %s_body<false, -1>(P, e, P.t, P.ldt, e, P.p, P.ldp, P.out, P.ldo, k * P.ldk);" % name)
    src.append("}")
    if True:    # bilinear forms and load vectors alike (a vector = one CSR slot per row)
        nve = 4 if quad else dim + 1
        # element id (only needed for the interp gathers through P.tdof_u)
        elem = "(long long)((const int *)(sA + d[13]))[el]" if spec.needs_el else "0LL"
        cases = "\n".join(
            "        case %d: %s_body<true, %d>(T.base, %s, vt, (long long)ne_pad, (long long)el, "
            "pc, ldpc, sL, (long long)ldl, (long long)el); break;" % (c, name, c, elem)
            for c in range(spec.npart))
        pcases = "\n".join(
            "            case %d: %s_body<true, %d>(T.base, %s, vt, (long long)ne_pad, (long long)el, "
            "pc, ldpc, po, 1LL, 0LL); break;" % (c, name, c, elem)
            for c in range(spec.npart))
        src.append(TILE_KERNEL.replace("@CASES@", cases).replace("@PCASES@", pcases)
                   .replace("@UT@", ", ".join(str(u) for u in ut))
                   .replace("@NAME@", name)
                   .replace("@BUNROLL@", os.environ.get("SPFEM_TILE_BUNROLL", "2"))
                   .replace("@TUNROLL@", os.environ.get("SPFEM_TILE_TUNROLL", "4"))
                   .replace("@THREADS@", str(TILE_THREADS))
                   # registers: three CTAs of 256 threads per SM on 2-D meshes (<= 85
                   # registers; measured +19 % for vector-P2), two on 3-D meshes
                   # (the quadrature loop of the larger quadrilateral elements keeps ~25
                   # accumulators + the H vectors live: 128 registers, two CTAs)
                   .replace("@MINB@", "SPFEM_TILE_MINB"))
    src.append("#else")
    src.append('extern "C" void %s_host(const SpfemArgs *P)\n{' % name)
    src.append("    for (long long k = 0; k < P->n; ++k) {")
    src.append("        const long long e = P->sel ? (long long)P->sel[k] : k;")
    src.append("        %s_body<false, -1>(*P, e, P->t, P->ldt, e, P->p, P->ldp, P->out, P->ldo, k * P->ldk);" % name)
    src.append("    }")
    src.append("}")
    src.append("#endif")
    return "\n".join(src) + "\n"


# ---------------------------------------------------------------------------
# facet kernels
# ---------------------------------------------------------------------------
class FacetSpec(object):
    """``fasm`` kernel description.  ``blocks``: list over the facet blocks --
    one for boundary facets, four ((1,1),(2,2),(2,1),(1,2) = (side_u, side_v),
    spfem/assembly.py:1131-1149) for interior facets -- of
    ``{(cu, cv): Form}``."""

    def __init__(self, refdom, dim, bilinear, interior, nbu, nbv, ncu, ncv,
                 blocks, sides, X, W, polys_u, polys_v, normals=True, nw=0,
                 functional=False, nfld=0):
        # functional: one value per facet (fnorm) -- no test functions, and the
        # interpolated fields exist on both sides of an interior facet
        self.functional = functional
        self.refdom, self.dim, self.bilinear = refdom, dim, bilinear
        self.interior = interior
        self.nbu, self.nbv, self.ncu, self.ncv = nbu, nbv, ncu, ncv
        self.blocks, self.sides, self.X, self.W = blocks, sides, X, W
        self.polys_u, self.polys_v = polys_u, polys_v
        self.normals, self.nw = normals, nw
        self.nfld = nfld            # host-evaluated coefficient fields, [q * ldf + facet]

    @property
    def n_rows(self):
        return self.nbv * self.ncv

    @property
    def n_cols(self):
        return self.nbu * self.ncu if self.bilinear else 1

    @property
    def n_entries(self):
        return self.n_rows * self.n_cols


_NREF = {2: ((0.0, -1.0), (1.0, 1.0), (-1.0, 0.0)),
         3: ((0.0, 0.0, -1.0), (0.0, -1.0, 0.0), (-1.0, 0.0, 0.0),
             (1.0, 1.0, 1.0))}


def generate_facet(spec, name="spfem_facet"):
    """Source text of the facet kernel: one thread per facet, run-time loops
    over quadrature points and basis functions."""
    dim, nq = spec.dim, len(spec.W)
    bd = dim - 1                       # dimension of the reference facet
    quad = spec.refdom == "quad"
    all_forms = [f for blk in spec.blocks for f in blk.values()]
    syms = collect_syms(all_forms)
    nsides = 2 if spec.interior else 1
    NA = 1 + dim
    L = []
    A = L.append
    A("const long long f = k;")
    for s in range(1, nsides + 1):
        A("const long long el%d = P.e%d[f];" % (s, s))
    # facet geometry: B, c, detB            (spfem/mapping.py:234-242, 284-302)
    for i in range(dim):
        A("const double c_%d = P.p[%d * P.ldp + P.fv[f]];" % (i, i))
        for j in range(bd):
            A("const double B_%d%d = %s;" % (
                i, j, _sub("P.p[%d * P.ldp + P.fv[%d * P.n + f]]" % (i, j + 1),
                           "c_%d" % i)))
    if dim == 2:
        A("const double detB = sqrt(%s);" % _add(_mul("B_00", "B_00"),
                                                 _mul("B_10", "B_10")))
    else:
        A("const double cr0 = %s;" % _sub(_mul("B_10", "B_21"), _mul("B_20", "B_11")))
        A("const double cr1 = %s;" % _add(_mul(_neg("B_00"), "B_21"), _mul("B_20", "B_01")))
        A("const double cr2 = %s;" % _sub(_mul("B_00", "B_11"), _mul("B_10", "B_01")))
        A("const double detB = sqrt(%s);" % _add(_add(_mul("cr0", "cr0"),
                                                      _mul("cr1", "cr1")),
                                                 _mul("cr2", "cr2")))
    A("const double absdetB = fabs(detB);")
    A("const double hf = pow(absdetB, %s);" % lit(1.0 / (dim - 1.0)))
    # element geometry on each side
    for s in range(1, nsides + 1):
        if quad:
            # MappingQ1.invF = one Newton step from the origin
            #                                   (spfem/mapping.py:124-139)
            for kk in range(4):
                A("const int vq%d_%d = P.t[%d * P.ldt + el%d];" % (s, kk, kk, s))
            for i in range(2):
                for kk in range(4):
                    A("const double Q%d_%d = P.p[%d * P.ldp + vq%d_%d];"
                      % (i, kk, i, s, kk))
            for ln in q1_geometry_q("c%d" % s, 0.0, 0.0):
                A(ln)
            for i in range(2):
                acc = _mul("Q%d_0" % i, lit(_q1_basis(0.0, 0.0, 0)))
                for c in range(1, 4):
                    acc = _add(acc, _mul("Q%d_%d" % (i, c), lit(_q1_basis(0.0, 0.0, c))))
                A("const double F0%d_%d = %s;" % (s, i, acc))
        else:
            for ln in affine_geometry(dim, sfx=str(s), elem="el%d" % s):
                A(ln)
    # outward normal of side 1           (spfem/mapping.py:474-524)
    if spec.normals and not quad:
        A("const int lf = P.lf1[f];")
        for k in range(dim):
            sel = "0.0"
            for itr in range(len(_NREF[dim]) - 1, -1, -1):
                sel = "(lf == %d ? %s : %s)" % (itr, lit(_NREF[dim][itr][k]), sel)
            A("const double nr_%d = %s;" % (k, sel))
        for k in range(dim):
            acc = _mul("iA1_0%d" % k, "nr_0")
            for l in range(1, dim):
                acc = _add(acc, _mul("iA1_%d%d" % (l, k), "nr_%d" % l))
            A("const double nn_%d = %s;" % (k, acc))
        acc = _mul("nn_0", "nn_0")
        for k in range(1, dim):
            acc = _add(acc, _mul("nn_%d" % k, "nn_%d" % k))
        A("const double nlen = sqrt(%s);" % acc)
        for k in range(dim):
            A("const double n%d = %s;" % (k, _div("nn_%d" % k, "nlen")))
        if dim == 2 and ("t0" in syms or "t1" in syms):
            # tangent t = (-n_1, n_0)            (spfem/assembly.py:1239-1241)
            A("const double t0 = -n1;")
            A("const double t1 = n0;")
    A("double acc[%d];" % (len(spec.blocks) * spec.n_entries))
    A("for (int z = 0; z < %d; ++z) acc[z] = 0.0;" % (len(spec.blocks) * spec.n_entries))
    A("const double XQ[%d][%d] = {%s};" % (
        max(bd, 1), nq, ", ".join("{" + ", ".join(lit(spec.X[r, q]) for q in range(nq)) + "}"
                                  for r in range(bd))))
    A("const double WQ[%d] = {%s};" % (nq, ", ".join(lit(w) for w in spec.W)))
    A("#pragma unroll 1")
    A("for (int q = 0; q < %d; ++q) {" % nq)
    # x = G(X)                               (spfem/mapping.py:389-415)
    for i in range(dim):
        acc = _mul("B_%d0" % i, "XQ[0][q]")
        for j in range(1, bd):
            acc = _add(acc, _mul("B_%d%d" % (i, j), "XQ[%d][q]" % j))
        A("  const double x%d = %s;" % (i, _add(acc, "c_%d" % i)))
    A("  const double wdet = WQ[q] * absdetB;")
    for i in range(getattr(spec, "nfld", 0)):
        A("  const double fld%d = P.fields[%d][(long long)q * P.ldf + k];" % (i, i))
    for s in range(1, nsides + 1):
        # Y = invF(x)                         (spfem/mapping.py:350-387)
        if quad:
            for i in range(2):
                A("  const double g%d_%d = %s;" % (s, i, _sub("x%d" % i, "F0%d_%d" % (s, i))))
            for i in range(2):
                A("  const double Y%d_%d = %s;" % (s, i, _add(_add("0.0", _mul(
                    "iJ_%d0_c%d" % (i, s), "g%d_0" % s)), _mul("iJ_%d1_c%d" % (i, s), "g%d_1" % s))))
        else:
            for i in range(dim):
                A("  const double g%d_%d = %s;" % (s, i, _sub("x%d" % i, "b%d_%d" % (s, i))))
            for i in range(dim):
                acc = _mul("iA%d_%d0" % (s, i), "g%d_0" % s)
                for j in range(1, dim):
                    acc = _add(acc, _mul("iA%d_%d%d" % (s, i, j), "g%d_%d" % (s, j)))
                A("  const double Y%d_%d = %s;" % (s, i, acc))
        ynames = ["Y%d_%d" % (s, i) for i in range(dim)]
        if quad:
            # per-point inverse Jacobian at Y (gbasis -> invDF(Y), element.py:474)
            A("  double iJ%d[2][2];" % s)
            A("  {")
            xq, yq = ynames
            for i in range(2):
                P_ = ["Q%d_%d" % (i, kk) for kk in range(4)]
                if s == 2:
                    pass
                e0 = _sub(_add(_add(_mul(_neg(P_[0]), _sub("1.0", yq)), _mul(P_[1], _sub("1.0", yq))),
                               _mul(P_[2], _add("1.0", yq))), _mul(P_[3], _add("1.0", yq)))
                e1 = _add(_add(_sub(_mul(_neg(P_[0]), _sub("1.0", xq)), _mul(P_[1], _add("1.0", xq))),
                               _mul(P_[2], _add("1.0", xq))), _mul(P_[3], _sub("1.0", xq)))
                A("    const double j%d0 = %s;" % (i, _mul("0.25", e0)))
                A("    const double j%d1 = %s;" % (i, _mul("0.25", e1)))
            A("    const double dj = %s;" % _sub(_mul("j00", "j11"), _mul("j01", "j10")))
            A("    iJ%d[0][0] = j11 / dj; iJ%d[0][1] = -j01 / dj;" % (s, s))
            A("    iJ%d[1][0] = -j10 / dj; iJ%d[1][1] = j00 / dj;" % (s, s))
            A("  }")
        for side, (polys, nb, tag) in enumerate(((spec.polys_u, spec.nbu, "pu"),
                                                 (spec.polys_v, spec.nbv, "pv"))):
            if side == 0 and not spec.bilinear and spec.nw == 0:
                continue
            if side == 1 and getattr(spec, "functional", False):
                continue
            P_, dP_ = polys
            A("  double %s%d[%d][%d];" % (tag, s, NA, nb))
            for j in range(nb):
                A("  %s%d[0][%d] = %s;" % (tag, s, j, P_[j].c_expr(ynames)))
                for l in range(dim):
                    A("  %s%d[%d][%d] = %s;" % (tag, s, 1 + l, j,
                                                dP_[j][l].c_expr(ynames)))
    # interpolated w / dw on side 1           (spfem/assembly.py:1076-1089);
    # fnorm also interpolates on side 2 (w<k>b, dw<k>b_<a>; assembly.py:1244-1262)
    for sl in range(spec.nw):
        for s_, tag in ((1, ""), (2, "b")):
            if s_ > nsides:
                continue
            if not (("w%d%s" % (sl, tag)) in syms
                    or any(("dw%d%s_%d" % (sl, tag, a)) in syms for a in range(dim))):
                continue
            A("  double w%d%s = 0.0;" % (sl, tag))
            for a in range(dim):
                A("  double dw%d%s_%d = 0.0;" % (sl, tag, a))
            A("  for (int j = 0; j < %d; ++j) {" % spec.nbu)
            A("    const double cf = P.interp[%d][P.tdof_u[j * P.ldd + el%d]];" % (sl, s_))
            A("    w%d%s += cf * pu%d[0][j];" % (sl, tag, s_))
            for a in range(dim):
                if quad:
                    g = " + ".join("iJ%d[%d][%d] * pu%d[%d][j]" % (s_, l, a, s_, 1 + l)
                                   for l in range(dim))
                else:
                    g = " + ".join("iA%d_%d%d * pu%d[%d][j]" % (s_, l, a, s_, 1 + l)
                                   for l in range(dim))
                A("    dw%d%s_%d += cf * (%s);" % (sl, tag, a, g))
            A("  }")

    def symname(nm, scope):
        # iA1_lk / iA2_lk : inverse Jacobians of side 1 / 2
        if quad and nm.startswith("iA"):
            return "iJ%s[%s][%s]" % (nm[2], nm[4], nm[5])
        return nm

    em = Emitter()
    em_lines_start = len(L)
    for bi, blk in enumerate(spec.blocks):
        su, sv = spec.sides[bi]
        for (cu, cv), form in sorted(blk.items()):
            for (a, b), coef in sorted(form.terms.items(), key=lambda kv: kv[0]):
                for node in (coef,):
                    # q-dependence is irrelevant inside the run-time loop
                    pass
                before = len(em.lines)
                g = em.name(coef, "", symname)
                for ln in em.lines[before:]:
                    A("  " + ln.strip())
                del em.lines[before:]
                pu = "1.0" if a == NONE else "pu%d[%d][j]" % (su, a)
                pv = "1.0" if b == NONE else "pv%d[%d][i]" % (sv, b)
                nbu = spec.nbu if spec.bilinear else 1
                A("  for (int j = 0; j < %d; ++j)" % nbu)
                A("    for (int i = 0; i < %d; ++i)" % spec.nbv)
                A("      acc[%d + (j * %d + %d) * %d + (i * %d + %d)] += wdet * (%s) * %s * %s;"
                  % (bi * spec.n_entries, spec.ncu if spec.bilinear else 1, cu,
                     spec.n_rows, spec.ncv, cv, g, pu, pv))
    A("}")
    nblk = len(spec.blocks)
    A("for (int b = 0; b < %d; ++b)" % nblk)
    A("  for (int z = 0; z < %d; ++z)" % spec.n_entries)
    A("    P.out[(long long)z * P.ldo + (long long)b * P.n + k] = acc[b * %d + z];"
      % spec.n_entries)

    spec.params = list(em.params)
    src = [PRELUDE]
    src.append("#define SPFEM_NENTRIES %d" % spec.n_entries)
    src.append("SPFEM_DEVICE void %s_body(const SpfemArgs &P, const long long k)\n{" % name)
    src += ["    " + ln for ln in L]
    src.append("}")
    src.append("#ifndef SPFEM_HOST")
    src.append('extern "C" __global__ void __launch_bounds__(64) %s(const SpfemArgs P)\n{' % name)
    src.append("    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;")
    src.append("    if (k < P.n) %s_body(P, k);" % name)
    src.append("}")
    src.append("#else")
    src.append('extern "C" void %s_host(const SpfemArgs *P)\n{' % name)
    src.append("    for (long long k = 0; k < P->n; ++k) %s_body(*P, k);" % name)
    src.append("}")
    src.append("#endif")
    return "\n".join(src) + "\n"


def source_key(src):
    return hashlib.sha256(src.encode("utf-8")).hexdigest()[:24]
