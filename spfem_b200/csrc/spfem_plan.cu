// spfem_plan -- native builder of the sparsity pattern AND the tile plan of the
// fused assembly kernel (spfem_b200/codegen.py, TILE_KERNEL), see
// include/spfem_b200.h "tile plan".
//
// Replaces two things of round 1: the global (row, col) sort of all
// Nu*Nv*nt local entries (spfem_pattern_sort/fill; limited to 2^31-1 entries
// per call) and the PyTorch index-op tile planner (dozens of int64 temporaries
// of that size: out of memory at 8 M vector-P2 triangles).  Here the work is
// tile-local and batched, so memory is bounded by the batch size, not by the
// mesh:
//
//   1. every CSR row gets an owner tile: the tile of E Morton-consecutive elements
//      that holds most of the row's elements (one radix sort of (row, tile) keys),
//      or simply the lowest tile touching it (integer atomicMin),
//   2. tile element lists = own elements + halo (elements touching an owned
//      row): (tile, element) pairs, one radix sort,
//   3. batches of tiles with <= batch_cap contributions: the contributions
//      (tile, owned row, column) -> shared-memory cell of the local value are
//      generated, radix-sorted (stable: fixed summation order) and run-length
//      encoded PER BATCH; this yields at the same time
//        * the CSR structure of the owned rows (= the rows of SciPy's
//          coo_matrix(...).tocsr(), spfem/assembly.py:985-986: sorted unique
//          columns, explicit zeros kept), and
//        * the per-tile slot / contribution lists of the kernel's phase B,
//   4. a prefix sum over the row lengths gives indptr; the column ids are
//      scattered from tile-major to CSR order and the row positions are patched
//      into the blobs.
//
// Vector elements (ElementH1Vec; DOF = bs * scalar DOF + component, spfem/
// assembly.py:1499-1550, spfem/element.py:505-546) are planned on the SCALAR
// structure: one "slot" stands for a bsv x bsu block of CSR values, which
// divides the plan's size and the kernel's index work by bsv*bsu.
//
// Everything runs on the device (CUB sorts / scans + small kernels) or, with
// on_device = 0, on the host with the same code (tests without a GPU: the
// lambdas are __host__ __device__).  Memory comes from the caller's allocator
// callbacks (PyTorch tensors in the Python host layer): no hidden allocation.
#include "../../include/spfem_b200.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

void spfem_set_error_(const std::string &msg);   // spfem_b200.cu

namespace spfem_plan {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef long long i64;

struct Err { int code; std::string msg; };

#define PLAN_CK(expr)                                                           \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess)                                                  \
            throw Err{1, std::string(#expr) + ": " + cudaGetErrorString(_e)};   \
    } while (0)

struct Ctx {
    bool dev;
    cudaStream_t st;
    spfem_alloc_fn alloc;
    spfem_free_fn release;
    void *actx;
    std::vector<void *> scratch;
    void *tmp = nullptr;          // CUB temporary storage, grown on demand
    size_t tmp_bytes = 0;
};

void *raw_alloc(Ctx &c, size_t bytes, const char *tag)
{
    void *p = c.alloc(c.actx, bytes ? bytes : 16, tag);
    if (!p) throw Err{3, std::string("allocation failed: ") + tag};
    return p;
}

template <class T> T *scratch(Ctx &c, size_t n, const char *tag)
{
    void *p = raw_alloc(c, n * sizeof(T), tag);
    c.scratch.push_back(p);
    return (T *)p;
}

template <class T> T *keep(Ctx &c, size_t n, const char *tag)
{
    return (T *)raw_alloc(c, n * sizeof(T), tag);
}

void drop(Ctx &c, void *p)
{
    if (!p) return;
    for (size_t i = 0; i < c.scratch.size(); ++i)
        if (c.scratch[i] == p) {
            c.scratch.erase(c.scratch.begin() + i);
            c.release(c.actx, p);
            return;
        }
}

void drop_all(Ctx &c)
{
    if (c.dev) cudaStreamSynchronize(c.st);
    for (void *p : c.scratch) c.release(c.actx, p);
    c.scratch.clear();
    c.tmp = nullptr;
    c.tmp_bytes = 0;
}

void *cub_tmp(Ctx &c, size_t bytes)
{
    if (bytes > c.tmp_bytes) {
        if (c.tmp) { PLAN_CK(cudaStreamSynchronize(c.st)); drop(c, c.tmp); }
        c.tmp_bytes = bytes + (bytes >> 2) + 256;
        c.tmp = scratch<char>(c, c.tmp_bytes, "cub");
    }
    return c.tmp;
}

template <class F> __global__ void k_for(i64 n, F f)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

template <class F> void par_for(Ctx &c, i64 n, F f)
{
    if (n <= 0) return;
    if (c.dev) {
        k_for<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(n, f);
        PLAN_CK(cudaGetLastError());
    } else {
        for (i64 i = 0; i < n; ++i) f(i);
    }
}

__host__ __device__ inline void amin(int *p, int v)
{
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}

__host__ __device__ inline void amin64(long long *p, long long v)
{
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}

__host__ __device__ inline void amax(int *p, int v)
{
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}

__host__ __device__ inline int popc(u32 x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

template <class T> T get(Ctx &c, const T *p)
{
    T v;
    if (c.dev) {
        PLAN_CK(cudaMemcpyAsync(&v, p, sizeof(T), cudaMemcpyDeviceToHost, c.st));
        PLAN_CK(cudaStreamSynchronize(c.st));
    } else {
        v = *p;
    }
    return v;
}

template <class T> void to_host(Ctx &c, const T *p, size_t n, std::vector<T> &out)
{
    out.resize(n);
    if (!n) return;
    if (c.dev) {
        PLAN_CK(cudaMemcpyAsync(out.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, c.st));
        PLAN_CK(cudaStreamSynchronize(c.st));
    } else {
        memcpy(out.data(), p, n * sizeof(T));
    }
}

void from_host(Ctx &c, void *dst, const void *src, size_t bytes)
{
    if (!bytes) return;
    if (c.dev) {
        PLAN_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.st));
        PLAN_CK(cudaStreamSynchronize(c.st));
    } else {
        memcpy(dst, src, bytes);
    }
}

template <class T> void fill(Ctx &c, T *p, i64 n, T v)
{
    par_for(c, n, [=] __host__ __device__(i64 i) { p[i] = v; });
}

// out[i] = sum_{k<i} in[k] for i in [0, n]  (out has n + 1 entries)
void exscan(Ctx &c, const i64 *in, i64 *out, i64 n)
{
    if (c.dev) {
        if (n > 0) {
            size_t bytes = 0;
            PLAN_CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c.st));
            void *t = cub_tmp(c, bytes);
            PLAN_CK(cub::DeviceScan::ExclusiveSum(t, bytes, in, out, n, c.st));
        }
        par_for(c, 1, [=] __host__ __device__(i64) {
            out[n] = n > 0 ? out[n - 1] + in[n - 1] : 0;
        });
    } else {
        i64 acc = 0;
        for (i64 i = 0; i < n; ++i) { out[i] = acc; acc += in[i]; }
        out[n] = acc;
    }
}

int bits_for(u64 maxval)
{
    int b = 1;
    while (b < 64 && (maxval >> b) != 0ull) ++b;
    return b;
}

// stable sort of (key, value) pairs by the low `end_bit` bits of the key; on
// return `keys` / `vals` point at the sorted data (one of the two buffers)
void sort_pairs(Ctx &c, u64 *&keys, u64 *&keys_alt, u32 *&vals, u32 *&vals_alt, i64 n, int end_bit)
{
    if (n <= 1) return;
    if (c.dev) {
        cub::DoubleBuffer<u64> dk(keys, keys_alt);
        cub::DoubleBuffer<u32> dv(vals, vals_alt);
        size_t bytes = 0;
        PLAN_CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, end_bit, c.st));
        void *t = cub_tmp(c, bytes);
        PLAN_CK(cub::DeviceRadixSort::SortPairs(t, bytes, dk, dv, n, 0, end_bit, c.st));
        keys = dk.Current(); keys_alt = dk.Alternate();
        vals = dv.Current(); vals_alt = dv.Alternate();
    } else {
        std::vector<i64> idx(n);
        for (i64 i = 0; i < n; ++i) idx[i] = i;
        const u64 *k = keys;
        std::stable_sort(idx.begin(), idx.end(), [k](i64 a, i64 b) { return k[a] < k[b]; });
        for (i64 i = 0; i < n; ++i) { keys_alt[i] = keys[idx[i]]; vals_alt[i] = vals[idx[i]]; }
        std::swap(keys, keys_alt);
        std::swap(vals, vals_alt);
    }
}

void sort_keys(Ctx &c, u64 *&keys, u64 *&keys_alt, i64 n, int end_bit)
{
    if (n <= 1) return;
    if (c.dev) {
        cub::DoubleBuffer<u64> dk(keys, keys_alt);
        size_t bytes = 0;
        PLAN_CK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, dk, n, 0, end_bit, c.st));
        void *t = cub_tmp(c, bytes);
        PLAN_CK(cub::DeviceRadixSort::SortKeys(t, bytes, dk, n, 0, end_bit, c.st));
        keys = dk.Current(); keys_alt = dk.Alternate();
    } else {
        std::sort(keys, keys + n);
    }
}

// seg[g] = first position p in [0, n) with group(p) >= g, for g in [0, ngroups];
// group() is non-decreasing over the positions
template <class G> void boundaries(Ctx &c, i64 n, i64 ngroups, i64 *seg, G group)
{
    if (n == 0) { fill<i64>(c, seg, ngroups + 1, 0); return; }
    par_for(c, n, [=] __host__ __device__(i64 p) {
        const i64 g = group(p);
        const i64 gp = p == 0 ? -1 : group(p - 1);
        for (i64 x = gp + 1; x <= g; ++x) seg[x] = p;
        if (p == n - 1)
            for (i64 x = g + 1; x <= ngroups; ++x) seg[x] = n;
    });
}

__host__ __device__ inline i64 pad16(i64 x) { return (x + 15) & ~(i64)15; }

struct Geo {                  // scalar structure of the DOF tables
    const int *rowdof; i64 ldr; int nbv, bsv;
    const int *coldof; i64 ldc; int nbu, bsu;
    __host__ __device__ int row(int i, i64 e) const { return rowdof[(i64)(bsv * i) * ldr + e] / bsv; }
    __host__ __device__ int col(int j, i64 e) const { return coldof ? coldof[(i64)(bsu * j) * ldc + e] / bsu : 0; }
};

#define MAXLOC 64             // scalar local DOFs per element (rows)

int build(const SpfemPlanIn &in, SpfemPlanOut &out, Ctx &c)
{
    memset(&out, 0, sizeof(out));
    const i64 n = in.n;
    if (n <= 0) throw Err{1, "tile plan: no elements"};
    if (in.E <= 0 || in.E > 32767) throw Err{1, "tile plan: bad E"};
    if (in.nrows <= 0 || in.nrows > INT_MAX || in.ncols > INT_MAX)
        throw Err{1, "tile plan: row / column count out of range"};
    int bsv = in.bsv > 0 ? in.bsv : 1, bsu = in.bsu > 0 ? in.bsu : 1;
    if (!in.coldof) bsu = 1;
    if (in.nloc_v % bsv || in.nrows % bsv) bsv = 1;
    if (in.coldof && (in.nloc_u % bsu || in.ncols % bsu)) bsu = 1;
    if (in.compact) { bsv = 1; bsu = 1; }

    // ---- 0. block structure of the DOF tables ------------------------------
    if (bsv > 1 || bsu > 1) {
        int *flag = scratch<int>(c, 2, "flag");
        fill<int>(c, flag, 2, 1);
        const int *rd = in.rowdof, *cd = in.coldof;
        const i64 ldr = in.ldr, ldc = in.ldc;
        const int nlv = in.nloc_v, nlu = in.nloc_u, bv = bsv, bu = bsu;
        par_for(c, n, [=] __host__ __device__(i64 e) {
            bool okv = true, oku = true;
            for (int i = 0; i < nlv; i += bv) {
                const int d0 = rd[(i64)i * ldr + e];
                if (d0 % bv) okv = false;
                for (int a = 1; a < bv; ++a)
                    if (rd[(i64)(i + a) * ldr + e] != d0 + a) okv = false;
            }
            if (cd)
                for (int j = 0; j < nlu; j += bu) {
                    const int d0 = cd[(i64)j * ldc + e];
                    if (d0 % bu) oku = false;
                    for (int b = 1; b < bu; ++b)
                        if (cd[(i64)(j + b) * ldc + e] != d0 + b) oku = false;
                }
            if (!okv) flag[0] = 0;
            if (!oku) flag[1] = 0;
        });
        if (!get(c, flag)) bsv = 1;
        if (!get(c, flag + 1)) bsu = 1;
        drop(c, flag);
    }
    Geo g;
    g.rowdof = in.rowdof; g.ldr = in.ldr; g.bsv = bsv; g.nbv = in.nloc_v / bsv;
    g.coldof = in.coldof; g.ldc = in.ldc; g.bsu = bsu; g.nbu = in.coldof ? in.nloc_u / bsu : 1;
    const int nbv = g.nbv, nbu = g.nbu, nb = bsv * bsu;
    if (nbv > MAXLOC) throw Err{1, "tile plan: more than 64 scalar test functions per element"};
    const i64 NR = in.nrows / bsv;
    const i64 NC = in.coldof ? in.ncols / bsu : 1;
    const int E = in.E;
    i64 ntiles = (n + E - 1) / E;
    const unsigned char *row_mask = in.row_mask;
    const int INF = INT_MAX;

    // ---- 1 + 2. tiles, row owners, tile element lists (own elements + halo) ----
    // Tiles are E consecutive elements; a tile whose element list (with halo) is
    // far above the median -- the Morton curve jumps at quadrant boundaries, so a
    // few tiles span distant regions -- is split, because the largest tile sizes
    // the shared memory of every CTA.
    std::vector<i64> h_tstart;
    for (i64 e0 = 0; e0 < n; e0 += E) h_tstart.push_back(e0);
    h_tstart.push_back(n);
    int *etile = scratch<int>(c, n, "etile");
    int *owner = scratch<int>(c, NR, "owner");
    int *pe = nullptr, *pt = nullptr;
    i64 *te_ptr = nullptr;
    i64 P = 0;
    for (int pass = 0;; ++pass) {
        ntiles = (i64)h_tstart.size() - 1;
        i64 *tstart = scratch<i64>(c, ntiles + 1, "tstart");
        from_host(c, tstart, h_tstart.data(), (size_t)(ntiles + 1) * sizeof(i64));
        {
            const i64 nt_ = ntiles;
            par_for(c, n, [=] __host__ __device__(i64 e) {
                i64 lo = 0, hi = nt_;
                while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (tstart[m] <= e) lo = m; else hi = m; }
                etile[e] = (int)lo;
            });
        }
        drop(c, tstart);
        fill<int>(c, owner, NR, INF);
        if (in.owner_rule == 0) {
            // lowest tile touching the row
            par_for(c, n * nbv, [=] __host__ __device__(i64 w) {
                const int i = (int)(w / n);
                const i64 e = w - (i64)i * n;
                const int R = g.row(i, e);
                if (row_mask && !row_mask[(i64)R * bsv]) return;
                amin(&owner[R], etile[e]);
            });
        } else {
            // the tile holding MOST of the row's elements (ties: the lowest): fewer halo
            // elements, and above all balanced tiles -- with "lowest tile" the first tile
            // of a neighbourhood owns every shared row (TetP2: up to 184 rows against a
            // mean of 90), and the largest tile sizes the shared memory of every CTA
            const i64 nW = n * nbv;
            const u64 nt64 = (u64)ntiles;
            u64 *rk = scratch<u64>(c, nW, "own:rk"), *rk2 = scratch<u64>(c, nW, "own:rk2");
            par_for(c, nW, [=] __host__ __device__(i64 w) {
                const int i = (int)(w / n);
                const i64 e = w - (i64)i * n;
                const int R = g.row(i, e);
                const bool off = row_mask && !row_mask[(i64)R * bsv];
                rk[w] = off ? ~(u64)0 : (u64)R * nt64 + (u64)etile[e];
            });
            sort_keys(c, rk, rk2, nW, 64);
            drop(c, rk2);
            u64 *best = scratch<u64>(c, NR, "own:best");
            fill<u64>(c, best, NR, (u64)0);
            {
                const u64 *k = rk;
                par_for(c, nW, [=] __host__ __device__(i64 w) {
                    const u64 key = k[w];
                    if (key == ~(u64)0 || (w > 0 && k[w - 1] == key)) return;     // heads of runs only
                    i64 lo = w, hi = nW;                                            // first index with k > key
                    while (lo < hi) { const i64 m = (lo + hi) >> 1; if (k[m] <= key) lo = m + 1; else hi = m; }
                    const u64 R = key / nt64, T = key - R * nt64;
                    const u64 packed = ((u64)(lo - w) << 32) | (u64)(u32)(nt64 - 1 - T);
#ifdef __CUDA_ARCH__
                    atomicMax((unsigned long long *)&best[R], (unsigned long long)packed);
#else
                    if (packed > best[R]) best[R] = packed;
#endif
                });
            }
            par_for(c, NR, [=] __host__ __device__(i64 R) {
                if (best[R] != 0) owner[R] = (int)(nt64 - 1 - (best[R] & 0xffffffffull));
            });
            drop(c, rk); drop(c, best);
        }
        i64 *cnt = scratch<i64>(c, n + 1, "cnt");
        par_for(c, n, [=] __host__ __device__(i64 e) {
            int seen[MAXLOC];
            int ns = 0;
            for (int i = 0; i < nbv; ++i) {
                const int o = owner[g.row(i, e)];
                if (o == INF) continue;
                bool dup = false;
                for (int k = 0; k < ns; ++k) dup = dup || seen[k] == o;
                if (!dup) seen[ns++] = o;
            }
            cnt[e] = ns;
        });
        i64 *poff = scratch<i64>(c, n + 1, "poff");
        exscan(c, cnt, poff, n);
        P = get(c, poff + n);
        if (P == 0) throw Err{1, "tile plan: no owned rows"};
        u64 *pk = scratch<u64>(c, P, "pairs"), *pk2 = scratch<u64>(c, P, "pairs2");
        par_for(c, n, [=] __host__ __device__(i64 e) {
            int seen[MAXLOC];
            int ns = 0;
            i64 pos = poff[e];
            for (int i = 0; i < nbv; ++i) {
                const int o = owner[g.row(i, e)];
                if (o == INF) continue;
                bool dup = false;
                for (int k = 0; k < ns; ++k) dup = dup || seen[k] == o;
                if (!dup) { seen[ns++] = o; pk[pos++] = (u64)o * (u64)n + (u64)e; }
            }
        });
        drop(c, cnt); drop(c, poff);
        sort_keys(c, pk, pk2, P, bits_for((u64)ntiles * (u64)n));
        drop(c, pk2);
        pe = scratch<int>(c, P, "pe");
        pt = scratch<int>(c, P, "pt");
        {
            const u64 *k = pk;
            int *pe_ = pe, *pt_ = pt;
            par_for(c, P, [=] __host__ __device__(i64 q) {
                const u64 T = k[q] / (u64)n;
                pt_[q] = (int)T;
                pe_[q] = (int)(k[q] - T * (u64)n);
            });
        }
        drop(c, pk);
        te_ptr = scratch<i64>(c, ntiles + 1, "te_ptr");
        {
            const int *pt_ = pt;
            boundaries(c, P, ntiles, te_ptr, [=] __host__ __device__(i64 q) { return (i64)pt_[q]; });
        }
        if ((in.split_factor <= 0.0 && in.ne_budget <= 0) || pass >= 4) break;
        std::vector<i64> h_te;
        to_host(c, te_ptr, (size_t)ntiles + 1, h_te);
        std::vector<i64> ne_sorted(ntiles);
        i64 ne_hi = 0;
        for (i64 T = 0; T < ntiles; ++T) {
            ne_sorted[T] = h_te[T + 1] - h_te[T];
            ne_hi = std::max(ne_hi, ne_sorted[T]);
        }
        std::nth_element(ne_sorted.begin(), ne_sorted.begin() + ntiles / 2, ne_sorted.end());
        const i64 median = ne_sorted[ntiles / 2];
        // the shared-memory image has one column per tile element, padded to a
        // multiple of 32: the budget is the padded size of a typical tile
        i64 budget = std::max<i64>((i64)(in.split_factor * (double)median + 0.999), 8);
        budget = (budget + 31) / 32 * 32;
        if (in.ne_budget > 0) budget = std::max<i64>(in.ne_budget, median + 1);
        if (ne_hi <= budget) break;
        std::vector<i64> nt2;
        for (i64 T = 0; T < ntiles; ++T) {
            const i64 m = h_tstart[T + 1] - h_tstart[T];
            const i64 ne = h_te[T + 1] - h_te[T];
            i64 parts = 1;
            if (ne > budget) parts = std::min<i64>(m, (ne * 5 + budget * 4 - 1) / (budget * 4));
            for (i64 k = 0; k < parts; ++k) nt2.push_back(h_tstart[T] + (m * k) / parts);
        }
        nt2.push_back(n);
        if (nt2.size() == h_tstart.size()) break;
        h_tstart.swap(nt2);
        drop(c, pe); drop(c, pt); drop(c, te_ptr);
    }
    drop(c, etile);
    int *mx = scratch<int>(c, 8, "mx");       // running maxima: ne, nr, nc, ns, nvt, store, bytes_a, bytes_b
    fill<int>(c, mx, 8, 0);
    par_for(c, ntiles, [=] __host__ __device__(i64 T) { amax(&mx[0], (int)(te_ptr[T + 1] - te_ptr[T])); });
    const int ne_max = get(c, mx);
    const int ldl = ne_max | 1;          // odd: consecutive value rows shift banks
    int els = 0;
    while ((1 << els) < ne_max) ++els;
    const int nsidx = nbv * nbu;
    const bool compact = in.compact != 0;
    if (nb > 1) {
        if (((i64)nsidx << els) > 65536) throw Err{2, "tile too large for 16-bit contribution codes (blocks)"};
    } else if (!compact) {
        if ((i64)in.nuniq * ldl > 65535) throw Err{2, "tile too large for 16-bit shared-memory indices"};
    }

    // ---- 3. contributions per (tile, element) pair ----------------------------
    i64 *cntc = scratch<i64>(c, P + 1, "cntc");
    par_for(c, P, [=] __host__ __device__(i64 q) {
        const int T = pt[q];
        const i64 e = pe[q];
        int k = 0;
        for (int i = 0; i < nbv; ++i) k += owner[g.row(i, e)] == T;
        cntc[q] = (i64)k * nbu;
    });
    i64 *coff = scratch<i64>(c, P + 1, "coff");
    exscan(c, cntc, coff, P);
    drop(c, cntc);
    const i64 n_contrib = get(c, coff + P);
    // compact mode: which distinct local values of a tile element are read at all
    const int W = (in.nuniq + 31) / 32;
    const int *alias = in.alias;
    u32 *mask = nullptr;
    i64 *sb = nullptr;
    if (compact) {
        mask = scratch<u32>(c, (size_t)P * W, "mask");
        i64 *pc = scratch<i64>(c, P + 1, "popc");
        par_for(c, P, [=] __host__ __device__(i64 q) {
            const int T = pt[q];
            const i64 e = pe[q];
            u32 m[32];
            for (int w = 0; w < W; ++w) m[w] = 0u;
            for (int i = 0; i < nbv; ++i) {
                if (owner[g.row(i, e)] != T) continue;
                for (int j = 0; j < nbu; ++j) {
                    const int u = alias[j * nbv + i];
                    m[u >> 5] |= 1u << (u & 31);
                }
            }
            i64 tot = 0;
            for (int w = 0; w < W; ++w) {
                mask[(size_t)q * W + w] = m[w];
                tot += popc(m[w]);
            }
            pc[q] = tot;
        });
        if (W > 32) throw Err{1, "tile plan: more than 1024 distinct local values"};
        sb = scratch<i64>(c, P + 1, "sbase");
        exscan(c, pc, sb, P);
        drop(c, pc);
        par_for(c, ntiles, [=] __host__ __device__(i64 T) {
            amax(&mx[5], (int)(sb[te_ptr[T + 1]] - sb[te_ptr[T]] > INT_MAX ? INT_MAX
                               : sb[te_ptr[T + 1]] - sb[te_ptr[T]]));
        });
        if (get(c, mx + 5) > 65535) throw Err{2, "tile too large for 16-bit shared-memory indices (compact)"};
    }

    // ---- 4. owned rows of every tile, rank of a row inside its tile ------------
    u64 *rk = scratch<u64>(c, NR, "rk"), *rk2 = scratch<u64>(c, NR, "rk2");
    par_for(c, NR, [=] __host__ __device__(i64 R) {
        const u64 o = owner[R] == INF ? (u64)ntiles : (u64)owner[R];
        rk[R] = o * (u64)NR + (u64)R;
    });
    sort_keys(c, rk, rk2, NR, bits_for((u64)(ntiles + 1) * (u64)NR));
    drop(c, rk2);
    i64 *tr_ptr = scratch<i64>(c, ntiles + 2, "tr_ptr");
    {
        const u64 *k = rk;
        boundaries(c, NR, ntiles + 1, tr_ptr, [=] __host__ __device__(i64 p) { return (i64)(k[p] / (u64)NR); });
    }
    int *rrank = scratch<int>(c, NR, "rrank");
    int *rows_sorted = scratch<int>(c, NR, "rows_sorted");
    {
        const u64 *k = rk;
        par_for(c, NR, [=] __host__ __device__(i64 p) {
            const u64 T = k[p] / (u64)NR;
            const int R = (int)(k[p] - T * (u64)NR);
            rows_sorted[p] = R;
            rrank[R] = (int)(p - tr_ptr[T]);
        });
    }
    drop(c, rk);
    par_for(c, ntiles, [=] __host__ __device__(i64 T) {
        amax(&mx[1], (int)(tr_ptr[T + 1] - tr_ptr[T]));
        const i64 ncT = coff[te_ptr[T + 1]] - coff[te_ptr[T]];
        amax(&mx[2], (int)(ncT > INT_MAX ? INT_MAX : ncT));
    });
    const int nr_max = get(c, mx + 1), nc_max = get(c, mx + 2);
    if (nr_max > 65535) throw Err{2, "more than 65535 owned rows in one tile"};
    if (nc_max > 65535) throw Err{2, "more than 65535 contributions in one tile"};
    const int cbits = bits_for((u64)(NC > 1 ? NC - 1 : 1));
    const int rbits = bits_for((u64)(nr_max > 1 ? nr_max - 1 : 1));
    const int tbits = 64 - cbits - rbits;
    if (tbits < 1) throw Err{1, "tile plan: key does not fit 64 bits"};
    const i64 max_tiles_batch = tbits >= 40 ? ((i64)1 << 40) : ((i64)1 << tbits);

    // ---- batches ----------------------------------------------------------------
    std::vector<i64> h_tc;                       // contributions before tile T
    {
        i64 *tc = scratch<i64>(c, ntiles + 1, "tc");
        par_for(c, ntiles + 1, [=] __host__ __device__(i64 T) { tc[T] = coff[te_ptr[T]]; });
        to_host(c, tc, (size_t)ntiles + 1, h_tc);
        drop(c, tc);
    }
    const i64 cap = in.batch_cap > 0 ? in.batch_cap : ((i64)1 << 27);
    std::vector<i64> bstart;
    for (i64 T = 0; T < ntiles;) {
        bstart.push_back(T);
        i64 T2 = T + 1;
        while (T2 < ntiles && T2 - T < max_tiles_batch && h_tc[T2 + 1] - h_tc[T] <= cap) ++T2;
        T = T2;
    }
    bstart.push_back(ntiles);
    const int nbatch = (int)bstart.size() - 1;
    if (nbatch > SPFEM_PLAN_MAX_BATCHES) throw Err{1, "tile plan: too many batches (raise batch_cap)"};

    const int sorted = in.sort_slots;
    const int nve = in.nve, dim = in.dim, with_el = in.with_el;
    const int *vt = in.vt;
    const i64 ldvt = in.ldvt;
    i64 *rowlen = scratch<i64>(c, NR + 1, "rowlen");
    fill<i64>(c, rowlen, NR + 1, 0);

    struct Pending {             // kept until indptr is known
        int *cols; i64 nS; i64 *rowfirst; i64 nRb; i64 Ta; i64 *blob_off; i64 *bytes_a;
        i64 *ts_ptr; int *srow;
    };
    std::vector<Pending> pend;

    for (int b = 0; b < nbatch; ++b) {
        const i64 Ta = bstart[b], Tb = bstart[b + 1], nT = Tb - Ta;
        const i64 qa = get(c, te_ptr + Ta), qb = get(c, te_ptr + Tb), nQ = qb - qa;
        const i64 ca = h_tc[Ta], nC = h_tc[Tb] - ca;
        const i64 ra = get(c, tr_ptr + Ta), nRb = get(c, tr_ptr + Tb) - ra;
        SpfemPlanBatch &B = out.batches[b];
        B.ntiles = nT;
        // -- contributions: key (tile, row rank, column), value = shared-memory cell
        u64 *ck = scratch<u64>(c, nC, "ck"), *ck2 = scratch<u64>(c, nC, "ck2");
        u32 *cv = scratch<u32>(c, nC, "cv"), *cv2 = scratch<u32>(c, nC, "cv2");
        par_for(c, nQ, [=] __host__ __device__(i64 qq) {
            const i64 q = qa + qq;
            const int T = pt[q];
            const i64 e = pe[q];
            const u32 el = (u32)(q - te_ptr[T]);
            i64 pos = coff[q] - ca;
            for (int i = 0; i < nbv; ++i) {
                const int R = g.row(i, e);
                if (owner[R] != T) continue;
                const u64 hi = ((u64)(T - Ta) << (rbits + cbits)) | ((u64)rrank[R] << cbits);
                for (int j = 0; j < nbu; ++j) {
                    const int sidx = j * nbv + i;
                    u32 code;
                    if (nb > 1) {
                        code = ((u32)sidx << els) | el;
                    } else if (compact) {
                        const int u = alias[sidx];
                        i64 below = sb[q] - sb[te_ptr[T]];
                        for (int w = 0; w < (u >> 5); ++w) below += popc(mask[(size_t)q * W + w]);
                        below += popc(mask[(size_t)q * W + (u >> 5)] & ((1u << (u & 31)) - 1u));
                        code = (u32)below;
                    } else {
                        code = (u32)alias[sidx] * (u32)ldl + el;
                    }
                    ck[pos] = hi | (u64)g.col(j, e);
                    cv[pos] = code;
                    ++pos;
                }
            }
        });
        sort_pairs(c, ck, ck2, cv, cv2, nC, rbits + cbits + bits_for((u64)(nT > 1 ? nT - 1 : 1)));
        drop(c, ck2); drop(c, cv2);
        // -- slots = runs of equal keys
        i64 *head = scratch<i64>(c, nC + 1, "head");
        {
            const u64 *k = ck;
            par_for(c, nC, [=] __host__ __device__(i64 p) { head[p] = (p == 0 || k[p] != k[p - 1]) ? 1 : 0; });
        }
        i64 *sid = scratch<i64>(c, nC + 1, "sid");
        exscan(c, head, sid, nC);
        const i64 nS = get(c, sid + nC);
        u64 *skey = scratch<u64>(c, nS, "skey");
        i64 *sfc = scratch<i64>(c, nS + 1, "sfc");        // first contribution of every slot
        {
            const u64 *k = ck;
            par_for(c, nC, [=] __host__ __device__(i64 p) {
                if (head[p]) { skey[sid[p]] = k[p]; sfc[sid[p]] = p; }
                if (p == nC - 1) sfc[nS] = nC;
            });
        }
        drop(c, head); drop(c, sid); drop(c, ck);
        // -- per tile / per row structure (slots are in (tile, row rank, column) order)
        i64 *ts_ptr = scratch<i64>(c, nT + 1, "ts_ptr");
        boundaries(c, nS, nT, ts_ptr, [=] __host__ __device__(i64 s) { return (i64)(skey[s] >> (rbits + cbits)); });
        i64 *rowfirst = scratch<i64>(c, nRb + 1, "rowfirst");
        {
            const u64 rmask = ((u64)1 << rbits) - 1;
            boundaries(c, nS, nRb, rowfirst, [=] __host__ __device__(i64 s) {
                const i64 T = (i64)(skey[s] >> (rbits + cbits));
                return tr_ptr[Ta + T] - ra + (i64)((skey[s] >> cbits) & rmask);
            });
        }
        par_for(c, nRb, [=] __host__ __device__(i64 r) {
            rowlen[rows_sorted[ra + r]] = rowfirst[r + 1] - rowfirst[r];
        });
        int *cols = scratch<int>(c, nS, "cols");
        {
            const u64 cmask = ((u64)1 << cbits) - 1;
            par_for(c, nS, [=] __host__ __device__(i64 s) { cols[s] = (int)(skey[s] & cmask); });
        }
        // -- order of a tile's slots in the kernel (CSR order, or by contribution count)
        i64 *order = nullptr;                     // kernel slot position -> slot
        if (sorted) {
            u64 *ok = scratch<u64>(c, nS, "ok"), *ok2 = scratch<u64>(c, nS, "ok2");
            u32 *ov = scratch<u32>(c, nS, "ov"), *ov2 = scratch<u32>(c, nS, "ov2");
            if (nS > (i64)UINT_MAX) throw Err{1, "tile plan: batch has more than 2^32 slots"};
            par_for(c, nS, [=] __host__ __device__(i64 s) {
                const u64 T = skey[s] >> (rbits + cbits);
                i64 cn = sfc[s + 1] - sfc[s];
                if (cn > 65535) cn = 65535;
                u64 cls;
                if (sorted == 2) cls = 3 - ((cn > 2) + (cn > 4) + (cn > 8));
                else cls = 65535 - (u64)cn;
                ok[s] = (T << 16) | cls;
                ov[s] = (u32)s;
            });
            sort_pairs(c, ok, ok2, ov, ov2, nS, 16 + bits_for((u64)(nT > 1 ? nT - 1 : 1)));
            order = scratch<i64>(c, nS, "order");
            {
                const u32 *v = ov;
                par_for(c, nS, [=] __host__ __device__(i64 s) { order[s] = (i64)v[s]; });
            }
            drop(c, ok); drop(c, ok2); drop(c, ov); drop(c, ov2);
        }
        // start of every kernel slot's contributions inside its tile
        i64 *kcnt = scratch<i64>(c, nS + 1, "kcnt");
        par_for(c, nS, [=] __host__ __device__(i64 s) {
            const i64 so = order ? order[s] : s;
            kcnt[s] = sfc[so + 1] - sfc[so];
        });
        i64 *kofs = scratch<i64>(c, nS + 1, "kofs");
        exscan(c, kcnt, kofs, nS);
        drop(c, kcnt);
        // -- tile-local vertex numbering: unique (tile, vertex) keys
        const i64 nVK = nQ * nve;
        u64 *vk = scratch<u64>(c, nVK, "vk"), *vk2 = scratch<u64>(c, nVK, "vk2");
        par_for(c, nVK, [=] __host__ __device__(i64 w) {
            const int r = (int)(w / nQ);
            const i64 q = qa + (w - (i64)r * nQ);
            vk[w] = ((u64)(pt[q] - Ta) << 32) | (u64)(u32)vt[(i64)r * ldvt + pe[q]];
        });
        sort_keys(c, vk, vk2, nVK, 32 + bits_for((u64)(nT > 1 ? nT - 1 : 1)));
        drop(c, vk2);
        i64 *vh = scratch<i64>(c, nVK + 1, "vh");
        {
            const u64 *k = vk;
            par_for(c, nVK, [=] __host__ __device__(i64 p) { vh[p] = (p == 0 || k[p] != k[p - 1]) ? 1 : 0; });
        }
        i64 *vidx = scratch<i64>(c, nVK + 1, "vidx");
        exscan(c, vh, vidx, nVK);
        const i64 nU = get(c, vidx + nVK);
        u64 *uv = scratch<u64>(c, nU, "uv");
        {
            const u64 *k = vk;
            par_for(c, nVK, [=] __host__ __device__(i64 p) { if (vh[p]) uv[vidx[p]] = k[p]; });
        }
        drop(c, vh); drop(c, vidx); drop(c, vk);
        i64 *vptr = keep<i64>(c, nT + 1, "out:vptr");
        boundaries(c, nU, nT, vptr, [=] __host__ __device__(i64 p) { return (i64)(uv[p] >> 32); });
        int *vid = keep<int>(c, nU, "out:vid");
        par_for(c, nU, [=] __host__ __device__(i64 p) { vid[p] = (int)(uv[p] & 0xffffffffull); });
        // -- section sizes, blob offsets, descriptors
        i64 *tbytes = scratch<i64>(c, nT + 1, "tbytes");
        i64 *bytes_a = scratch<i64>(c, nT, "bytes_a");
        int *desc = keep<int>(c, 16 * nT, "out:desc");
        par_for(c, nT, [=] __host__ __device__(i64 t) {
            const i64 T = Ta + t;
            const i64 ne = te_ptr[T + 1] - te_ptr[T], ne_pad = (ne + 7) / 8 * 8;
            const i64 nvt = vptr[t + 1] - vptr[t], nvt_pad = (nvt + 1) / 2 * 2;
            const i64 nr = tr_ptr[T + 1] - tr_ptr[T];
            const i64 ns = ts_ptr[t + 1] - ts_ptr[t];
            const i64 nc = kofs[ts_ptr[t + 1]] - kofs[ts_ptr[t]];
            const i64 off_el = 2 * nve * ne_pad;
            const i64 off_pc = pad16(off_el + (with_el ? 4 * ne_pad : 0));
            i64 ba = off_pc + 8 * dim * nvt_pad;
            i64 off_base = 0;
            if (compact) { off_base = ba; ba = pad16(off_base + (1 + W) * 4 * ne_pad); }
            ba = pad16(ba);
            // section B: base position of the tile (int64); then
            //   slots ordered by contribution count: per slot the 32-bit offset of its
            //   (first) CSR value from the base [+ the length of its scalar row for block
            //   plans];
            //   slots in CSR order: one 8-byte record per ROW (offset of the row, first
            //   slot, length) and one 8-byte header per 32 slots (bit l: slot l starts a
            //   new row; row of slot 0) -- 0.25 bytes per slot instead of 4 or 6;
            // the start of every slot's contributions, the contribution lists
            const i64 off_pos = 16;
            const i64 off_slen = off_pos + (sorted ? pad16(4 * ns) : pad16(8 * nr));
            const i64 off_soff = off_slen + (sorted ? (nb > 1 ? pad16(2 * ns) : 0)
                                                    : pad16(8 * ((ns + 31) / 32)));
            const i64 off_contrib = off_soff + pad16(2 * (ns + 1));
            const i64 bb = off_contrib + pad16(2 * (nc + 1));
            bytes_a[t] = ba;
            tbytes[t] = ba + bb;
            int *d = desc + 16 * t;
            d[1] = (int)ba; d[2] = (int)bb; d[3] = (int)ne; d[4] = (int)ne_pad; d[5] = (int)ns;
            d[6] = (int)off_soff; d[7] = (int)off_slen; d[8] = (int)off_contrib; d[9] = (int)nr;
            d[10] = (int)nc; d[11] = (int)nvt_pad; d[12] = (int)off_pc;
            d[13] = with_el ? (int)off_el : -1; d[14] = (int)nvt; d[15] = (int)off_base;
            amax(&mx[3], (int)(ns > INT_MAX ? INT_MAX : ns));
            amax(&mx[4], (int)(nvt > INT_MAX ? INT_MAX : nvt));
            amax(&mx[6], (int)ba);
            amax(&mx[7], (int)(bb > INT_MAX ? INT_MAX : bb));
        });
        i64 *blob_off = scratch<i64>(c, nT + 1, "blob_off");
        exscan(c, tbytes, blob_off, nT);
        drop(c, tbytes);
        const i64 blob_bytes = get(c, blob_off + nT);
        if (get(c, mx + 3) > 65535) throw Err{2, "more than 65535 slots in one tile"};
        if (get(c, mx + 4) > 65535) throw Err{2, "more than 65535 vertices in one tile"};
        if (blob_bytes / 16 > INT_MAX) throw Err{1, "tile plan: batch blob exceeds 32 GB (lower batch_cap)"};
        unsigned char *blob = keep<unsigned char>(c, blob_bytes, "out:blob");
        if (c.dev) PLAN_CK(cudaMemsetAsync(blob, 0, blob_bytes, c.st));
        else memset(blob, 0, blob_bytes);
        par_for(c, nT, [=] __host__ __device__(i64 t) { desc[16 * t] = (int)(blob_off[t] / 16); });
        // -- section A: tile-local vertex ids, element ids, packed-value masks
        par_for(c, nQ, [=] __host__ __device__(i64 qq) {
            const i64 q = qa + qq;
            const i64 T = pt[q], t = T - Ta;
            const i64 e = pe[q];
            const i64 el = q - te_ptr[T];
            const int *d = desc + 16 * t;
            unsigned char *A = blob + blob_off[t];
            u16 *lv = (u16 *)A;
            const i64 ne_pad = d[4];
            const u64 *seg = uv + vptr[t];
            const i64 nvt = vptr[t + 1] - vptr[t];
            for (int r = 0; r < nve; ++r) {
                const u64 key = ((u64)t << 32) | (u64)(u32)vt[(i64)r * ldvt + e];
                i64 lo = 0, hi = nvt;
                while (lo < hi) { const i64 m = (lo + hi) >> 1; if (seg[m] < key) lo = m + 1; else hi = m; }
                lv[(i64)r * ne_pad + el] = (u16)lo;
            }
            if (with_el) ((int *)(A + d[13]))[el] = (int)e;
            if (compact) {
                int *base = (int *)(A + d[15]);
                base[el] = (int)(sb[q] - sb[te_ptr[T]]);
                u32 *mw = (u32 *)(base + ne_pad);
                for (int w = 0; w < W; ++w) mw[(i64)w * ne_pad + el] = mask[(size_t)q * W + w];
            }
        });
        drop(c, uv);
        // -- section B: slots, contributions (CSR positions are patched in step 6)
        int *srow_keep = scratch<int>(c, nS > 0 ? nS : 1, "srow_keep");
        par_for(c, nS, [=] __host__ __device__(i64 s) {
            const i64 so = order ? order[s] : s;
            const i64 t = (i64)(skey[so] >> (rbits + cbits));
            const int *d = desc + 16 * t;
            unsigned char *Bs = blob + blob_off[t] + d[1];
            const i64 sl = s - ts_ptr[t];
            const i64 ns = d[5];
            const u64 rmask = ((u64)1 << rbits) - 1;
            const i64 rr = (i64)((skey[so] >> cbits) & rmask);
            const i64 rb = tr_ptr[Ta + t] - ra + rr;
            u16 *soff = (u16 *)(Bs + d[6]);
            const i64 k0 = kofs[s] - kofs[ts_ptr[t]];
            soff[sl] = (u16)k0;
            if (sl == ns - 1) soff[ns] = (u16)d[10];
            srow_keep[s] = (int)rb;
            if (sorted) ((u32 *)(Bs + 16))[sl] = (u32)(so - rowfirst[rb]);   // position inside the row, for now
            u16 *cb = (u16 *)(Bs + d[8]);
            const i64 c0 = sfc[so], c1 = sfc[so + 1];
            for (i64 k = c0; k < c1; ++k) cb[k0 + (k - c0)] = (u16)cv[k];
        });
        drop(c, cv); drop(c, skey); drop(c, sfc); drop(c, kofs);
        if (order) drop(c, order);
        B.desc = desc; B.blob = blob; B.vid = vid; B.vptr = vptr;
        B.blob_bytes = blob_bytes; B.nvid = nU; B.first_tile = Ta; B.nslots = nS;
        pend.push_back(Pending{cols, nS, rowfirst, nRb, Ta, blob_off, bytes_a, ts_ptr, srow_keep});
    }

    // ---- 6. indptr, indices, row positions in the blobs ---------------------------
    i64 *iptr = scratch<i64>(c, NR + 1, "iptr");
    exscan(c, rowlen, iptr, NR);
    const i64 nnz_s = get(c, iptr + NR);
    const i64 nnz = nnz_s * nb;
    const int ib = in.index_bytes == 8 ? 8 : 4;
    if (ib == 4 && nnz > INT_MAX) throw Err{1, "tile plan: nnz exceeds int32 (index_bytes must be 8)"};
    void *indptr = raw_alloc(c, (size_t)(in.nrows + 1) * ib, "out:indptr");
    void *indices = in.coldof ? raw_alloc(c, (size_t)(nnz > 0 ? nnz : 1) * ib, "out:indices") : nullptr;
    const i64 nrows_full = in.nrows;
    par_for(c, NR + 1, [=] __host__ __device__(i64 R) {
        if (R == NR) {
            if (ib == 4) ((int *)indptr)[nrows_full] = (int)nnz; else ((i64 *)indptr)[nrows_full] = nnz;
            return;
        }
        const i64 L = rowlen[R];
        for (int a = 0; a < bsv; ++a) {
            const i64 v = nb * iptr[R] + (i64)a * bsu * L;
            if (ib == 4) ((int *)indptr)[R * bsv + a] = (int)v; else ((i64 *)indptr)[R * bsv + a] = v;
        }
    });
    for (int b = 0; b < nbatch; ++b) {
        const Pending &pd = pend[b];
        const SpfemPlanBatch &B = out.batches[b];
        const i64 Ta = pd.Ta, Tb = Ta + B.ntiles;
        const i64 ra = get(c, tr_ptr + Ta);
        const int *desc = (const int *)B.desc;
        unsigned char *blob = (unsigned char *)B.blob;
        const i64 *rowfirst = pd.rowfirst, *blob_off = pd.blob_off;
        const int *cols = pd.cols;
        const i64 nRb = pd.nRb;
        // base position of every tile = the lowest CSR position of its rows; the slots
        // hold 32-bit offsets from it (one look-up per slot instead of slot -> row ->
        // row base / row info)
        const i64 nT = B.ntiles, nS = pd.nS;
        const i64 *ts_ptr = pd.ts_ptr;
        const int *srow_keep = pd.srow;
        i64 *tmin = scratch<i64>(c, nT + 1, "tmin");
        fill<i64>(c, tmin, nT + 1, LLONG_MAX);
        par_for(c, nRb, [=] __host__ __device__(i64 r) {
            i64 lo = Ta, hi = Tb;
            while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (tr_ptr[m] - ra <= r) lo = m; else hi = m; }
            amin64(&tmin[lo - Ta], nb * iptr[rows_sorted[ra + r]]);
        });
        par_for(c, nT, [=] __host__ __device__(i64 t) {
            if (desc[16 * t + 5] > 0)
                *(i64 *)(blob + blob_off[t] + desc[16 * t + 1]) = tmin[t];
        });
        if (!sorted) {
            // row records and chunk headers (slots of a tile: row after row, CSR order)
            par_for(c, nRb, [=] __host__ __device__(i64 r) {
                i64 lo = Ta, hi = Tb;
                while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (tr_ptr[m] - ra <= r) lo = m; else hi = m; }
                const i64 t = lo - Ta;
                const int *d = desc + 16 * t;
                if (d[5] <= 0) return;
                const int R = rows_sorted[ra + r];
                const i64 g = nb * iptr[R] - tmin[t];
                if (g > (i64)UINT_MAX || rowlen[R] > 65535) amax(&mx[2], INT_MAX);
                u32 *rec = (u32 *)(blob + blob_off[t] + d[1] + 16) + 2 * (r - (tr_ptr[lo] - ra));
                rec[0] = (u32)g;
                rec[1] = (u32)(rowfirst[r] - ts_ptr[t]) | ((u32)rowlen[R] << 16);
            });
            par_for(c, nS, [=] __host__ __device__(i64 s) {
                i64 lo = 0, hi = nT;
                while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (ts_ptr[m] <= s) lo = m; else hi = m; }
                const i64 sl = s - ts_ptr[lo];
                if (sl & 31) return;
                const int *d = desc + 16 * lo;
                u32 *hd = (u32 *)(blob + blob_off[lo] + d[1] + d[7]) + 2 * (sl >> 5);
                const i64 send = ts_ptr[lo + 1];
                u32 m = 0;
                for (int l = 1; l < 32 && s + l < send; ++l) {
                    const int step = srow_keep[s + l] - srow_keep[s + l - 1];
                    if (step) m |= (u32)1 << l;
                    if (step > 1 || step < 0) amax(&mx[2], INT_MAX - 1);      // (cannot happen: every row has a slot)
                }
                hd[0] = m;
                hd[1] = (u32)(srow_keep[s] - (tr_ptr[Ta + lo] - ra));
            });
        }
        else par_for(c, nS, [=] __host__ __device__(i64 s) {
            i64 lo = 0, hi = nT;
            while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (ts_ptr[m] <= s) lo = m; else hi = m; }
            const int *d = desc + 16 * lo;
            unsigned char *Bs = blob + blob_off[lo] + d[1];
            const i64 sl = s - ts_ptr[lo];
            const int R = rows_sorted[ra + srow_keep[s]];
            u32 *pos = (u32 *)(Bs + 16);
            const i64 g = nb * iptr[R] + (i64)bsu * pos[sl] - tmin[lo];
            if (g > (i64)UINT_MAX) amax(&mx[2], INT_MAX);
            pos[sl] = (u32)g;
            if (nb > 1) {
                if (rowlen[R] > 65535) amax(&mx[2], INT_MAX);
                ((u16 *)(Bs + d[7]))[sl] = (u16)rowlen[R];
            }
        });
        if (get(c, mx + 2) == INT_MAX - 1) throw Err{1, "tile plan: a row of a tile has no slot"};
        if (get(c, mx + 2) == INT_MAX)
            throw Err{1, "tile plan: the rows of one tile span more than 2^32 CSR positions (or a row has more than 65535 blocks)"};
        drop(c, tmin);
        par_for(c, nRb, [=] __host__ __device__(i64 r) {
            i64 lo = Ta, hi = Tb;
            while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (tr_ptr[m] - ra <= r) lo = m; else hi = m; }
            const int R = rows_sorted[ra + r];
            const i64 base = nb * iptr[R];
            if (!indices) return;
            const i64 L = rowfirst[r + 1] - rowfirst[r];
            for (i64 k = 0; k < L; ++k) {
                const i64 C = cols[rowfirst[r] + k];
                for (int a = 0; a < bsv; ++a)
                    for (int bb = 0; bb < bsu; ++bb) {
                        const i64 pos = base + (i64)a * bsu * L + bsu * k + bb;
                        if (ib == 4) ((int *)indices)[pos] = (int)(C * bsu + bb);
                        else ((i64 *)indices)[pos] = C * bsu + bb;
                    }
            }
        });
    }
    out.nnz = nnz;
    out.indptr = indptr;
    out.indices = indices;
    out.nbatches = nbatch;
    out.ldl = ldl;
    out.els = els;
    out.bsv = bsv; out.bsu = bsu;
    out.ntiles = ntiles;
    out.n_pairs = P;
    out.n_contrib = n_contrib;
    out.max_ne = ne_max;
    out.max_nr = nr_max;
    out.max_ns = get(c, mx + 3);
    out.max_store = get(c, mx + 5);
    out.max_bytes_a = get(c, mx + 6);
    out.max_bytes_b = get(c, mx + 7);
    return 0;
}

} // namespace spfem_plan

extern "C" {

int spfem_tileplan_build(const SpfemPlanIn *in, SpfemPlanOut *out, spfem_alloc_fn alloc,
                         spfem_free_fn release, void *alloc_ctx, int on_device, void *stream)
{
    using namespace spfem_plan;
    if (!in || !out || !alloc || !release) {
        spfem_set_error_("spfem_tileplan_build: null argument");
        return 1;
    }
    Ctx c;
    c.dev = on_device != 0;
    c.st = (cudaStream_t)stream;
    c.alloc = alloc; c.release = release; c.actx = alloc_ctx;
    int rc = 0;
    try {
        rc = build(*in, *out, c);
    } catch (const Err &e) {
        spfem_set_error_(e.msg);
        rc = e.code;
    } catch (const std::exception &e) {
        spfem_set_error_(std::string("spfem_tileplan_build: ") + e.what());
        rc = 1;
    }
    drop_all(c);
    return rc;
}

int spfem_tileplan_fill_coords(const int *desc, unsigned char *blob, const int *vid,
                               const long long *vptr, long long ntiles, const double *p,
                               long long ldp, int dim, int on_device, void *stream)
{
    using namespace spfem_plan;
    Ctx c;
    c.dev = on_device != 0;
    c.st = (cudaStream_t)stream;
    c.alloc = nullptr; c.release = nullptr; c.actx = nullptr;
    if (ntiles <= 0) return 0;
    try {
        const i64 total = get(c, vptr + ntiles);
        // one thread per tile vertex; the tile is found by binary search in vptr
        par_for(c, total, [=] __host__ __device__(i64 w) {
            i64 lo = 0, hi = ntiles;
            while (hi - lo > 1) { const i64 m = (lo + hi) >> 1; if (vptr[m] <= w) lo = m; else hi = m; }
            const int *d = desc + 16 * lo;
            double *pc = (double *)(blob + (i64)d[0] * 16 + d[12]);
            const i64 vl = w - vptr[lo];
            const i64 v = vid[w];
            if (dim == 2) {
                pc[2 * vl] = p[v];
                pc[2 * vl + 1] = p[ldp + v];
            } else {
                for (int k = 0; k < dim; ++k) pc[(i64)k * d[11] + vl] = p[(i64)k * ldp + v];
            }
        });
    } catch (const Err &e) {
        spfem_set_error_(e.msg);
        return e.code;
    }
    return 0;
}

} // extern "C"

// ===========================================================================
// Vertex-fan plan (codegen.FAN_KERNEL): scalar P1 on triangles, one thread per
// CSR row = vertex.  Replaces the PyTorch index-op builder of round 1 (hundreds
// of eager kernels, 0.2 - 0.5 s for 8 M triangles) by a handful of kernels +
// three radix sorts; tests/fanplan_ref.py is the NumPy statement of the same
// plan (byte-identical blobs).
// ===========================================================================
namespace spfem_plan {

#define FAN_MAXN 12

// monotone surrogate of atan2(dy, dx) in (0, 4] (no transcendental: identical on
// host and device): third, fourth, first, second quadrant, then "exactly left"
__host__ __device__ inline double fan_angle_key(double dx, double dy)
{
    if (dy == 0.0 && dx < 0.0) return 4.0;
    if (dy < 0.0) return dx < 0.0 ? (-dy) / (-dx - dy) : 1.0 + dx / (dx - dy);
    return dx > 0.0 ? 2.0 + dy / (dx + dy) : 3.0 + (-dx) / (-dx + dy);
}

__host__ __device__ inline void amin_u64(u64 *p, u64 v)
{
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}

__host__ __device__ inline u64 dbits(double x)
{
    union { double d; u64 u; } c;
    c.d = x;
    return c.u;
}

// index of `key` in the sorted array k[0, n), or -1
__host__ __device__ inline i64 find_key(const u64 *k, i64 n, u64 key)
{
    i64 lo = 0, hi = n;
    while (lo < hi) { const i64 m = (lo + hi) >> 1; if (k[m] < key) lo = m + 1; else hi = m; }
    return (lo < n && k[lo] == key) ? lo : -1;
}

struct IdxArray {            // indptr / indices of either width
    const void *a; int ib;
    __host__ __device__ i64 operator()(i64 k) const
    {
        return ib == 4 ? (i64)((const int *)a)[k] : ((const i64 *)a)[k];
    }
};

struct Corners {             // corner cid = i * nt + T: vertex r, CCW successor a, predecessor b
    const int *t; i64 ldt, nt; const unsigned char *ccw;
    __host__ __device__ void operator()(i64 cid, int &r, int &a, int &b) const
    {
        const int i = (int)(cid / nt);
        const i64 T = cid - (i64)i * nt;
        const int s = ccw[T] ? 1 : 2;
        r = t[(i64)i * ldt + T];
        a = t[(i64)((i + s) % 3) * ldt + T];
        b = t[(i64)((i + 3 - s) % 3) * ldt + T];
    }
};

int build_fan(const SpfemFanPlanIn &in, SpfemFanPlanOut &out, Ctx &c)
{
    memset(&out, 0, sizeof(out));
    const i64 nv = in.nv, nt = in.nt, nc3 = 3 * nt;
    if (nv <= 0 || nt <= 0) throw Err{5, "fan plan: empty mesh"};
    if (nv > INT_MAX || nc3 > (i64)UINT_MAX) throw Err{5, "fan plan: mesh too large"};
    const int R = in.rows_per_tile;
    if (R <= 0 || R > 32767) throw Err{1, "fan plan: bad rows_per_tile"};
    const double *p = in.p;
    const i64 ldp = in.ldp, ldt = in.ldt;
    const int *t = in.t;
    const int ib = in.index_bytes == 8 ? 8 : 4;
    const void *indptr = in.indptr, *indices = in.indices;
    const IdxArray IP{indptr, ib}, IX{indices, ib};      // (plain functors: captured by the kernels' lambdas)
    int *flag = scratch<int>(c, 8, "fan:flag");   // 0 degenerate, 1 non-manifold, 2 valence, 3 pattern
    fill<int>(c, flag, 8, 0);

    // ---- 1. corners: c = i * nt + T, vertex r, CCW successor a, predecessor b ------
    unsigned char *ccw = scratch<unsigned char>(c, nt, "fan:ccw");
    par_for(c, nt, [=] __host__ __device__(i64 T) {
        const i64 v0 = t[T], v1 = t[ldt + T], v2 = t[2 * ldt + T];
        const double x0 = p[v0], y0 = p[ldp + v0];
        const double det = (p[v1] - x0) * (p[ldp + v2] - y0) - (p[v2] - x0) * (p[ldp + v1] - y0);
        if (det == 0.0) flag[0] = 1;
        ccw[T] = det > 0.0 ? 1 : 0;
    });
    const Corners corner{t, ldt, nt, ccw};
    // directed edges r -> a, sorted: the corner across an edge is a binary search
    u64 *ek = scratch<u64>(c, nc3, "fan:ek"), *ek2 = scratch<u64>(c, nc3, "fan:ek2");
    u32 *ev = scratch<u32>(c, nc3, "fan:ev"), *ev2 = scratch<u32>(c, nc3, "fan:ev2");
    par_for(c, nc3, [=] __host__ __device__(i64 cid) {
        int r, a, b;
        corner(cid, r, a, b);
        ek[cid] = (u64)r * (u64)nv + (u64)a;
        ev[cid] = (u32)cid;
    });
    sort_pairs(c, ek, ek2, ev, ev2, nc3, bits_for((u64)nv * (u64)nv));
    drop(c, ek2); drop(c, ev2);
    {
        const u64 *k = ek;
        par_for(c, nc3 - 1, [=] __host__ __device__(i64 q) { if (k[q] == k[q + 1]) flag[1] = 1; });
    }
    // ---- 2. start corner of every vertex: the one without predecessor (boundary),
    //         else the smallest pseudo-angle of r -> a, ties to the smallest corner ----
    u64 *amin_key = scratch<u64>(c, nv, "fan:amin");
    int *valence = scratch<int>(c, nv, "fan:valence");
    fill<u64>(c, amin_key, nv, ~(u64)0);
    fill<int>(c, valence, nv, 0);
    double *ckey = scratch<double>(c, nc3, "fan:ckey");
    {
        const u64 *k = ek;
        par_for(c, nc3, [=] __host__ __device__(i64 cid) {
            int r, a, b;
            corner(cid, r, a, b);
            const bool has_prev = find_key(k, nc3, (u64)a * (u64)nv + (u64)r) >= 0;
            const double key = has_prev ? fan_angle_key(p[a] - p[r], p[ldp + a] - p[ldp + r]) : 0.0;
            ckey[cid] = key;
            amin_u64(&amin_key[r], dbits(key));
#ifdef __CUDA_ARCH__
            atomicAdd(&valence[r], 1);
#else
            valence[r] += 1;
#endif
        });
    }
    u64 *start = scratch<u64>(c, nv, "fan:start");
    fill<u64>(c, start, nv, ~(u64)0);
    par_for(c, nc3, [=] __host__ __device__(i64 cid) {
        int r, a, b;
        corner(cid, r, a, b);
        if (dbits(ckey[cid]) == amin_key[r]) amin_u64(&start[r], (u64)cid);
    });
    drop(c, ckey); drop(c, amin_key);
    // ---- 3. ring walk: one thread per vertex -----------------------------------------
    int *nb = scratch<int>(c, (size_t)FAN_MAXN * nv, "fan:nb");     // nb[k * nv + v]
    int *ring = scratch<int>(c, nv, "fan:ring");                    // nfan | closed << 8
    {
        const u64 *k = ek;
        const u32 *val = ev;
        par_for(c, nv, [=] __host__ __device__(i64 v) {
            for (int q = 0; q < FAN_MAXN; ++q) nb[(i64)q * nv + v] = -1;
            if (valence[v] == 0) { flag[2] = 1; ring[v] = 0; return; }
            const i64 st = (i64)start[v];
            i64 cur = st;
            int nfan = 0, closed = 0, last_b = -1;
            while (true) {
                int r, a, b;
                corner(cur, r, a, b);
                if (nfan >= FAN_MAXN) { flag[2] = 1; break; }
                nb[(i64)nfan * nv + v] = a;
                last_b = b;
                ++nfan;
                const i64 q = find_key(k, nc3, (u64)r * (u64)nv + (u64)b);   // next triangle CCW
                if (q < 0) break;
                const i64 nxt = (i64)val[q];
                if (nxt == st) { closed = 1; break; }
                cur = nxt;
            }
            if (nfan != valence[v]) flag[1] = 1;
            if (!closed) {
                if (nfan >= FAN_MAXN) flag[2] = 1;
                else nb[(i64)nfan * nv + v] = last_b;
            }
            ring[v] = nfan | (closed << 8);
        });
    }
    drop(c, ek); drop(c, ev); drop(c, start); drop(c, valence); drop(c, ccw);
    {
        std::vector<int> hf;
        to_host(c, flag, 8, hf);
        if (hf[0]) throw Err{5, "fan plan: degenerate triangle"};
        if (hf[2]) throw Err{5, "fan plan: vertex without elements or with more than 12 ring neighbours"};
        if (hf[1]) throw Err{5, "fan plan: non-manifold vertex fan"};
    }
    // ---- 4. positions of the ring neighbours inside the CSR rows ------------------------
    {
        const i64 nnz = ib == 4 ? (i64)get(c, (const int *)indptr + nv) : get(c, (const i64 *)indptr + nv);
        if (nnz > (i64)INT_MAX) throw Err{5, "fan plan: nnz exceeds 32-bit positions"};
    }
    unsigned char *pos = scratch<unsigned char>(c, (size_t)(FAN_MAXN + 1) * nv, "fan:pos");  // [k*nv+v], k = MAXN: diagonal
    int *mxn = scratch<int>(c, 4, "fan:mx");       // max ring neighbours, max nvt, max ns, max bytes
    fill<int>(c, mxn, 4, 0);
    par_for(c, nv, [=] __host__ __device__(i64 v) {
        const int nfan = ring[v] & 0xff, closed = ring[v] >> 8;
        const int nn = nfan + (closed ? 0 : 1);
        amax(&mxn[0], nn);
        const i64 r0 = IP(v), r1 = IP(v + 1);
        if (r1 - r0 != nn + 1 || r1 - r0 > 255) { flag[3] = 1; return; }
        for (int q = 0; q <= FAN_MAXN; ++q) {
            const i64 col = q == FAN_MAXN ? v : (q < nn ? (i64)nb[(i64)q * nv + v] : -1);
            if (col < 0) { pos[(i64)q * nv + v] = 0; continue; }
            i64 lo = r0, hi = r1;
            while (lo < hi) { const i64 m = (lo + hi) >> 1; if (IX(m) < col) lo = m + 1; else hi = m; }
            if (lo >= r1 || IX(lo) != col) { flag[3] = 1; return; }
            pos[(i64)q * nv + v] = (unsigned char)(lo - r0);
        }
    });
    if (get(c, flag + 3)) throw Err{5, "fan plan: pattern rows are not vertex stars"};
    const int max_nn = get(c, mxn);
    const bool compact = max_nn <= 8;
    const int W = compact ? 8 : FAN_MAXN;
    // ---- 5. tiles: Morton order of the vertices, rows of a tile in CSR (id) order --------
    const i64 ntiles = (nv + R - 1) / R;
    u64 *mk = scratch<u64>(c, nv, "fan:mk"), *mk2 = scratch<u64>(c, nv, "fan:mk2");
    u32 *mv = scratch<u32>(c, nv, "fan:mv"), *mv2 = scratch<u32>(c, nv, "fan:mv2");
    {
        const double lo0 = in.lo[0], lo1 = in.lo[1];
        const double e0 = in.hi[0] - lo0 == 0.0 ? 1.0 : in.hi[0] - lo0;
        const double e1 = in.hi[1] - lo1 == 0.0 ? 1.0 : in.hi[1] - lo1;
        par_for(c, nv, [=] __host__ __device__(i64 v) {
            const double s = 1048575.0;                 // 2^20 - 1
            i64 q0 = (i64)((p[v] - lo0) / e0 * s), q1 = (i64)((p[ldp + v] - lo1) / e1 * s);
            q0 = q0 < 0 ? 0 : (q0 > 1048575 ? 1048575 : q0);
            q1 = q1 < 0 ? 0 : (q1 > 1048575 ? 1048575 : q1);
            u64 key = 0;
            for (int b = 0; b < 20; ++b)
                key |= (((u64)q0 >> b) & 1) << (2 * b) | (((u64)q1 >> b) & 1) << (2 * b + 1);
            mk[v] = key;
            mv[v] = (u32)v;
        });
    }
    sort_pairs(c, mk, mk2, mv, mv2, nv, 40);
    if (in.row_order == 0) {
        const u32 *val = mv;
        u64 *k = mk;
        par_for(c, nv, [=] __host__ __device__(i64 j) { k[j] = (u64)(j / R) * (u64)nv + (u64)val[j]; });
        sort_pairs(c, mk, mk2, mv, mv2, nv, bits_for((u64)ntiles * (u64)nv));
    }
    drop(c, mk); drop(c, mk2); drop(c, mv2);
    const u32 *rows = mv;                            // vertex of tile-row position j
    int *pos_of = scratch<int>(c, nv, "fan:pos_of"); // tile-row position of a vertex
    par_for(c, nv, [=] __host__ __device__(i64 j) { pos_of[rows[j]] = (int)j; });
    // ---- 6. halo vertices: ring neighbours owned by another tile, unique per tile --------
    i64 *hcnt = scratch<i64>(c, nv + 1, "fan:hcnt");
    par_for(c, nv, [=] __host__ __device__(i64 j) {
        const i64 v = rows[j], T = j / R;
        int k = 0;
        for (int q = 0; q < W; ++q) {
            const int n = nb[(i64)q * nv + v];
            if (n >= 0 && pos_of[n] / R != T) ++k;
        }
        hcnt[j] = k;
    });
    i64 *hoff = scratch<i64>(c, nv + 1, "fan:hoff");
    exscan(c, hcnt, hoff, nv);
    drop(c, hcnt);
    const i64 nH = get(c, hoff + nv);
    u64 *hk = scratch<u64>(c, nH > 0 ? nH : 1, "fan:hk"), *hk2 = scratch<u64>(c, nH > 0 ? nH : 1, "fan:hk2");
    par_for(c, nv, [=] __host__ __device__(i64 j) {
        const i64 v = rows[j], T = j / R;
        i64 o = hoff[j];
        for (int q = 0; q < W; ++q) {
            const int n = nb[(i64)q * nv + v];
            if (n >= 0 && pos_of[n] / R != T) hk[o++] = (u64)T * (u64)nv + (u64)n;
        }
    });
    drop(c, hoff);
    sort_keys(c, hk, hk2, nH, bits_for((u64)ntiles * (u64)nv));
    drop(c, hk2);
    i64 *hh = scratch<i64>(c, nH + 1, "fan:hh");
    {
        const u64 *k = hk;
        par_for(c, nH, [=] __host__ __device__(i64 q) { hh[q] = (q == 0 || k[q] != k[q - 1]) ? 1 : 0; });
    }
    i64 *hid = scratch<i64>(c, nH + 1, "fan:hid");
    exscan(c, hh, hid, nH);
    const i64 nU = get(c, hid + nH);
    u64 *uh = scratch<u64>(c, nU > 0 ? nU : 1, "fan:uh");
    {
        const u64 *k = hk;
        par_for(c, nH, [=] __host__ __device__(i64 q) { if (hh[q]) uh[hid[q]] = k[q]; });
    }
    drop(c, hh); drop(c, hid); drop(c, hk);
    i64 *h0 = scratch<i64>(c, ntiles + 1, "fan:h0");
    boundaries(c, nU, ntiles, h0, [=] __host__ __device__(i64 q) { return (i64)(uh[q] / (u64)nv); });
    // ---- 7. per tile sizes, offsets, descriptors -----------------------------------------
    i64 *cl = scratch<i64>(c, nv + 1, "fan:cl");     // slots before tile-row position j
    {
        i64 *len = scratch<i64>(c, nv + 1, "fan:len");
        par_for(c, nv, [=] __host__ __device__(i64 j) { const i64 v = rows[j]; len[j] = IP(v + 1) - IP(v); });
        exscan(c, len, cl, nv);
        drop(c, len);
    }
    par_for(c, ntiles, [=] __host__ __device__(i64 T) {
        const i64 nr = (T + 1) * R <= nv ? R : nv - T * R;
        const i64 nvt = nr + h0[T + 1] - h0[T];
        amax(&mxn[1], (int)(nvt > INT_MAX ? INT_MAX : nvt));
    });
    const int max_nvt = get(c, mxn + 1);
    if (max_nvt > 65535) throw Err{5, "fan plan: tile with more than 65535 vertices"};
    const bool tiny = in.tiny != 0 && compact && max_nvt <= 256;
    const int nplanes = tiny ? 1 : (compact ? 2 : 3);
    i64 *tb = scratch<i64>(c, ntiles + 1, "fan:tb");
    int *desc = keep<int>(c, 8 * ntiles, "out:fan_desc");
    par_for(c, ntiles, [=] __host__ __device__(i64 T) {
        const i64 nr = (T + 1) * R <= nv ? R : nv - T * R;
        const i64 nvt = nr + h0[T + 1] - h0[T];
        const i64 ba = nvt * 16;
        const i64 bb = nr * 16 * nplanes + (tiny ? pad16(nr * 4) : 0);
        const i64 j0 = T * R;
        const i64 ns = cl[j0 + nr] - cl[j0];
        tb[T] = ba + bb;
        int *d = desc + 8 * T;
        d[1] = (int)ba; d[2] = (int)bb; d[3] = (int)nr; d[4] = (int)ns; d[5] = (int)nvt; d[6] = 0; d[7] = 0;
        amax(&mxn[2], (int)(ns > INT_MAX ? INT_MAX : ns));
        amax(&mxn[3], (int)(ba + bb));
    });
    i64 *off = scratch<i64>(c, ntiles + 1, "fan:off");
    exscan(c, tb, off, ntiles);
    drop(c, tb);
    const i64 total = get(c, off + ntiles);
    if (total / 16 > INT_MAX) throw Err{5, "fan plan offsets exceed 32 bits"};
    if (get(c, mxn + 2) > 65535) throw Err{5, "fan plan: more than 65535 slots in one tile"};
    par_for(c, ntiles, [=] __host__ __device__(i64 T) { desc[8 * T] = (int)(off[T] / 16); });
    unsigned char *blob = keep<unsigned char>(c, total > 0 ? total : 16, "out:fan_blob");
    if (c.dev) PLAN_CK(cudaMemsetAsync(blob, 0, total, c.st));
    else memset(blob, 0, total);
    // ---- 8. row records ---------------------------------------------------------------------
    par_for(c, nv, [=] __host__ __device__(i64 j) {
        const i64 v = rows[j], T = j / R, rl = j - T * R;
        const i64 nr = (T + 1) * R <= nv ? R : nv - T * R;
        const int nfan = ring[v] & 0xff, closed = ring[v] >> 8;
        const i64 nvt16 = (nr + h0[T + 1] - h0[T]) * 16;
        unsigned char *rec = blob + off[T] + nvt16;
        u32 loc[FAN_MAXN], ps[FAN_MAXN];
        for (int q = 0; q < FAN_MAXN; ++q) {
            loc[q] = 0; ps[q] = 0;
            if (q >= W) continue;
            const int n = nb[(i64)q * nv + v];
            if (n < 0) continue;
            ps[q] = pos[(i64)q * nv + v];
            const i64 pj = pos_of[n];
            if (pj / R == T) {
                loc[q] = (u32)(pj - T * R);
            } else {
                const i64 h = find_key(uh + h0[T], h0[T + 1] - h0[T], (u64)T * (u64)nv + (u64)n);
                loc[q] = (u32)(nr + h);
            }
        }
        const u32 w0 = (u32)IP(v);
        const u32 s0 = (u32)(cl[j] - cl[T * R]);
        const u32 w1 = s0 | ((u32)nfan << 16) | ((u32)closed << 20) | ((u32)pos[(i64)FAN_MAXN * nv + v] << 24);
        u32 *pl0 = (u32 *)(rec + 16 * rl);
        pl0[0] = w0; pl0[1] = w1;
        if (tiny) {
            pl0[2] = loc[0] | loc[1] << 8 | loc[2] << 16 | loc[3] << 24;
            pl0[3] = loc[4] | loc[5] << 8 | loc[6] << 16 | loc[7] << 24;
            u32 nib = 0;
            for (int q = 0; q < 8; ++q) nib |= ps[q] << (4 * q);
            ((u32 *)(rec + 16 * nr))[rl] = nib;
        } else {
            pl0[2] = loc[0] | loc[1] << 16;
            pl0[3] = loc[2] | loc[3] << 16;
            u32 *pl1 = (u32 *)(rec + 16 * nr + 16 * rl);
            pl1[0] = loc[4] | loc[5] << 16;
            pl1[1] = loc[6] | loc[7] << 16;
            if (compact) {
                pl1[2] = ps[0] | ps[1] << 8 | ps[2] << 16 | ps[3] << 24;
                pl1[3] = ps[4] | ps[5] << 8 | ps[6] << 16 | ps[7] << 24;
            } else {
                pl1[2] = loc[8] | loc[9] << 16;
                pl1[3] = loc[10] | loc[11] << 16;
                u32 *pl2 = (u32 *)(rec + 32 * nr + 16 * rl);
                pl2[0] = ps[0] | ps[1] << 8 | ps[2] << 16 | ps[3] << 24;
                pl2[1] = ps[4] | ps[5] << 8 | ps[6] << 16 | ps[7] << 24;
                pl2[2] = ps[8] | ps[9] << 8 | ps[10] << 16 | ps[11] << 24;
                pl2[3] = 0;
            }
        }
    });
    // ---- 9. coordinate slots: the rows first, then the halo vertices ---------------------------
    const i64 nvid = nv + nU;
    int *vid = keep<int>(c, nvid, "out:fan_vid");
    i64 *pcx = keep<i64>(c, nvid, "out:fan_pc_index");
    par_for(c, nvid, [=] __host__ __device__(i64 w) {
        if (w < nv) {
            const i64 T = w / R;
            vid[w] = (int)rows[w];
            pcx[w] = off[T] / 8 + 2 * (w - T * R);
        } else {
            const i64 q = w - nv;
            const i64 T = (i64)(uh[q] / (u64)nv);
            const i64 nr = (T + 1) * R <= nv ? R : nv - T * R;
            vid[w] = (int)(uh[q] - (u64)T * (u64)nv);
            pcx[w] = off[T] / 8 + 2 * (nr + q - h0[T]);
        }
    });
    out.desc = desc; out.blob = blob; out.blob_bytes = total;
    out.vid = vid; out.pc_index = pcx; out.nvid = nvid; out.ntiles = ntiles;
    out.max_a = get(c, mxn + 3);
    out.max_ns = get(c, mxn + 2);
    out.max_neighbours = max_nn;
    out.format = tiny ? 0 : (compact ? 1 : 2);
    out.halo_factor = (double)nvid / (double)nv;
    return 0;
}

} // namespace spfem_plan

extern "C" int spfem_fanplan_build(const SpfemFanPlanIn *in, SpfemFanPlanOut *out, spfem_alloc_fn alloc,
                                   spfem_free_fn release, void *alloc_ctx, int on_device, void *stream)
{
    using namespace spfem_plan;
    if (!in || !out || !alloc || !release) {
        spfem_set_error_("spfem_fanplan_build: null argument");
        return 1;
    }
    Ctx c;
    c.dev = on_device != 0;
    c.st = (cudaStream_t)stream;
    c.alloc = alloc; c.release = release; c.actx = alloc_ctx;
    int rc = 0;
    try {
        rc = build_fan(*in, *out, c);
    } catch (const Err &e) {
        spfem_set_error_(e.msg);
        rc = e.code;
    } catch (const std::exception &e) {
        spfem_set_error_(std::string("spfem_fanplan_build: ") + e.what());
        rc = 1;
    }
    drop_all(c);
    return rc;
}
