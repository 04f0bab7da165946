// spfem_b200 -- C-ABI library behind spfem_b200.assembly (see include/spfem_b200.h).
//
// Static kernels of the assembly path that do not depend on the weak form:
//   * sparsity pattern build: 64-bit (row, col) keys -> stable radix sort ->
//     run-length encoding -> CSR indptr / indices + contribution lists
//     (replaces SciPy's coo_tocsr + csr_sum_duplicates structure pass that
//     spfem/assembly.py:985-986 reaches through coo_matrix(...).tocsr()),
//   * deterministic duplicate summation (fixed-order gather, no fp atomics),
//   * Morton ordering of elements,
// plus the NVRTC front end and the loader / launcher of the generated element
// kernels (spfem_b200/codegen.py).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared
//        -Xcompiler -fPIC -cudart static  ... -lnvrtc
#include "../../include/spfem_b200.h"

#include <cuda_runtime.h>
#include <nvrtc.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_err;

int fail(const std::string &msg)
{
    g_err = msg;
    return 1;
}

#define CUDA_TRY(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess)                                                 \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- workspace layout of the pattern build ---------------------------------
struct PatternWs {
    unsigned long long *keys_a, *keys_b;
    int *ids_a, *ids_b;
    int *scan;          // inclusive scan of the run-head flags
    void *cub_tmp;
    size_t cub_bytes;
    size_t total;
};

size_t cub_tmp_bytes(long long n)
{
    size_t b1 = 0, b2 = 0;
    cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
    cub::DoubleBuffer<int> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b1, dk, dv, (int)n, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, b2, (int *)nullptr, (int *)nullptr, (int)n);
    return b1 > b2 ? b1 : b2;
}

PatternWs layout(void *base, long long n)
{
    PatternWs w;
    char *p = (char *)base;
    size_t off = 0;
    w.keys_a = (unsigned long long *)(p + off); off += align_up(8 * (size_t)n);
    w.keys_b = (unsigned long long *)(p + off); off += align_up(8 * (size_t)n);
    w.ids_a = (int *)(p + off); off += align_up(4 * (size_t)n);
    w.ids_b = (int *)(p + off); off += align_up(4 * (size_t)n);
    w.scan = (int *)(p + off); off += align_up(4 * (size_t)n);
    w.cub_bytes = align_up(cub_tmp_bytes(n));
    w.cub_tmp = p + off; off += w.cub_bytes;
    w.total = off;
    return w;
}

// ---- kernels -----------------------------------------------------------------
__global__ void k_make_keys(const int *__restrict__ rowdof, long long ldr,
                            const int *__restrict__ coldof, long long ldc,
                            long long n, int nloc_v, long long n_entries,
                            unsigned long long ncols,
                            unsigned long long *__restrict__ keys,
                            int *__restrict__ ids)
{
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_entries) return;
    long long blk = id / n;
    long long k = id - blk * n;
    int j = (int)(blk / nloc_v);
    int i = (int)(blk - (long long)j * nloc_v);
    unsigned long long r = (unsigned long long)rowdof[i * ldr + k];
    unsigned long long c = coldof ? (unsigned long long)coldof[j * ldc + k] : 0ull;
    keys[id] = r * ncols + c;
    ids[id] = (int)id;
}

__global__ void k_flag_heads(const unsigned long long *__restrict__ keys,
                             long long n, int *__restrict__ flags)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

template <typename I>
__global__ void k_fill_const(I *__restrict__ a, long long n, I v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// One thread per sorted entry.  Run heads write their CSR slot: column index,
// start of the contribution list, and the row pointers of every row that
// starts at this slot (rows without entries included).
template <typename I>
__global__ void k_fill_csr(const unsigned long long *__restrict__ keys,
                           const int *__restrict__ scan, long long n,
                           unsigned long long ncols, long long nnz,
                           I *__restrict__ indptr, I *__restrict__ indices,
                           int *__restrict__ cptr, int *__restrict__ rowstart)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key = keys[i];
    bool head = (i == 0) || (key != keys[i - 1]);
    if (!head) return;
    long long s = (long long)scan[i] - 1;
    long long r = (long long)(key / ncols);
    if (indices) indices[s] = (I)(key - (unsigned long long)r * ncols);
    if (cptr) cptr[s] = (int)i;
    long long rprev = (i == 0) ? -1 : (long long)(keys[i - 1] / ncols);
    if (r != rprev)
        for (long long rr = rprev + 1; rr <= r; ++rr) {
            if (indptr) indptr[rr] = (I)s;
            if (rowstart) rowstart[rr] = (int)i;
        }
    if (i == 0 && cptr) cptr[nnz] = (int)n;
}

// values[s] = sum of the local entries listed for slot s, in list order.
__global__ void __launch_bounds__(256)
k_reduce(const double *__restrict__ local, const int *__restrict__ cptr,
         const int *__restrict__ contrib, long long nslots,
         double *__restrict__ values)
{
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    int b = cptr[s], e = cptr[s + 1];
    double acc = 0.0;
    for (int k = b; k < e; ++k) acc += local[contrib[k]];
    values[s] = acc;
}

__device__ __forceinline__ unsigned long long spread2(unsigned long long x)
{   // 32 bits -> every second bit
    x &= 0xffffffffull;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long x)
{   // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

struct Box { double lo[3]; double inv[3]; };

__global__ void k_morton_keys(const double *__restrict__ p, long long ldp, int dim,
                              const int *__restrict__ t, long long ldt, int nve,
                              long long nt, Box box,
                              unsigned long long *__restrict__ keys,
                              int *__restrict__ ids)
{
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    double c[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < nve; ++k) {
        long long v = t[k * ldt + e];
        for (int d = 0; d < dim; ++d) c[d] += p[d * ldp + v];
    }
    unsigned long long key = 0;
    const double scale = (dim == 2) ? 4294967295.0 : 2097151.0;
    for (int d = 0; d < dim; ++d) {
        double u = (c[d] / nve - box.lo[d]) * box.inv[d];
        u = fmin(fmax(u, 0.0), 1.0);
        unsigned long long q = (unsigned long long)(u * scale);
        key |= (dim == 2 ? spread2(q) : spread3(q)) << d;
    }
    keys[e] = key;
    ids[e] = (int)e;
}


inline unsigned grid_for(long long n, int block)
{
    return (unsigned)((n + block - 1) / block);
}

int bits_for(unsigned long long maxval)
{
    int b = 1;
    while (b < 64 && (maxval >> b) != 0ull) ++b;
    return b;
}

} // namespace

void spfem_set_error_(const std::string &msg) { g_err = msg; }

extern "C" {

int spfem_abi_version(void) { return SPFEM_ABI_VERSION; }

const char *spfem_last_error(void) { return g_err.c_str(); }

int spfem_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void spfem_free(void *ptr) { free(ptr); }

int spfem_nvrtc_compile(const char *src, const char *arch, char **image,
                        size_t *image_size, char **log)
{
    if (image) *image = nullptr;
    if (log) *log = nullptr;
    if (!src || !image || !image_size) return fail("spfem_nvrtc_compile: null argument");
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, src, "spfem_kernel.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return fail(std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r));
    std::string archopt = std::string("--gpu-architecture=") + (arch && *arch ? arch : "sm_100a");
    const char *opts[] = {archopt.c_str(), "--std=c++17", "-lineinfo",
                          "--fmad=true", "-default-device"};
    r = nvrtcCompileProgram(prog, 5, opts);
    size_t logsz = 0;
    nvrtcGetProgramLogSize(prog, &logsz);
    std::string logstr;
    if (logsz > 1) {
        logstr.resize(logsz);
        nvrtcGetProgramLog(prog, &logstr[0]);
        if (log) {
            *log = (char *)malloc(logsz + 1);
            memcpy(*log, logstr.c_str(), logsz);
            (*log)[logsz] = 0;
        }
    }
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(std::string("nvrtcCompileProgram: ") + nvrtcGetErrorString(r) + "\n" + logstr);
    }
    size_t sz = 0;
    r = nvrtcGetCUBINSize(prog, &sz);
    if (r != NVRTC_SUCCESS || sz == 0) {
        nvrtcDestroyProgram(&prog);
        return fail("nvrtcGetCUBINSize failed (architecture must be a real sm_XX target)");
    }
    *image = (char *)malloc(sz);
    r = nvrtcGetCUBIN(prog, *image);
    nvrtcDestroyProgram(&prog);
    if (r != NVRTC_SUCCESS) {
        free(*image);
        *image = nullptr;
        return fail(std::string("nvrtcGetCUBIN: ") + nvrtcGetErrorString(r));
    }
    *image_size = sz;
    return 0;
}

int spfem_module_load(const void *image, size_t image_size, void **module)
{
    (void)image_size;
    if (!image || !module) return fail("spfem_module_load: null argument");
    cudaLibrary_t lib;
    CUDA_TRY(cudaLibraryLoadData(&lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0));
    *module = (void *)lib;
    return 0;
}

int spfem_module_unload(void *module)
{
    if (!module) return 0;
    CUDA_TRY(cudaLibraryUnload((cudaLibrary_t)module));
    return 0;
}

int spfem_module_get_kernel(void *module, const char *name, void **kernel)
{
    if (!module || !name || !kernel) return fail("spfem_module_get_kernel: null argument");
    cudaKernel_t k;
    CUDA_TRY(cudaLibraryGetKernel(&k, (cudaLibrary_t)module, name));
    *kernel = (void *)k;
    return 0;
}

int spfem_launch(void *kernel, const SpfemArgs *args, int block, void *stream)
{
    if (!kernel || !args) return fail("spfem_launch: null argument");
    if (args->n <= 0) return 0;
    if (block <= 0) block = 128;
    void *params[] = {(void *)args};
    CUDA_TRY(cudaLaunchKernel((const void *)kernel, dim3(grid_for(args->n, block)),
                              dim3(block), params, 0, (cudaStream_t)stream));
    return 0;
}

int spfem_launch_tiled(void *kernel, const void *args, unsigned grid, unsigned block,
                       size_t smem_bytes, void *stream)
{
    if (!kernel || !args) return fail("spfem_launch_tiled: null argument");
    if (grid == 0) return 0;
    if (smem_bytes > 48 * 1024) {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaKernelSetAttributeForDevice((cudaKernel_t)kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem_bytes, dev));
    }
    void *params[] = {(void *)args};
    CUDA_TRY(cudaLaunchKernel((const void *)kernel, dim3(grid), dim3(block), params,
                              smem_bytes, (cudaStream_t)stream));
    return 0;
}

size_t spfem_pattern_ws_bytes(long long n_entries)
{
    if (n_entries <= 0) n_entries = 1;
    return layout(nullptr, n_entries).total;
}

int spfem_pattern_sort(const int *d_rowdof, long long ldr,
                       const int *d_coldof, long long ldc,
                       long long n, int nloc_v, int nloc_u,
                       long long nrows, long long ncols,
                       void *d_ws, size_t ws_bytes, long long *h_nnz,
                       void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    long long n_entries = n * nloc_v * (d_coldof ? nloc_u : 1);
    if (!h_nnz) return fail("spfem_pattern_sort: null h_nnz");
    if (n_entries == 0) { *h_nnz = 0; return 0; }
    if (n_entries > 2147483647LL)
        return fail("spfem_pattern_sort: more than 2^31-1 local entries in one "
                    "call are not supported yet (shard the element range)");
    if (ncols <= 0 || nrows <= 0) return fail("spfem_pattern_sort: empty shape");
    PatternWs w = layout(d_ws, n_entries);
    if (ws_bytes < w.total) return fail("spfem_pattern_sort: workspace too small");
    const int B = 256;
    k_make_keys<<<grid_for(n_entries, B), B, 0, st>>>(d_rowdof, ldr, d_coldof, ldc, n,
                                                    nloc_v, n_entries,
                                                    (unsigned long long)ncols,
                                                    w.keys_a, w.ids_a);
    CUDA_TRY(cudaGetLastError());
    cub::DoubleBuffer<unsigned long long> dk(w.keys_a, w.keys_b);
    cub::DoubleBuffer<int> dv(w.ids_a, w.ids_b);
    int end_bit = bits_for((unsigned long long)nrows * (unsigned long long)ncols);
    size_t tmp = w.cub_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, dk, dv, (int)n_entries,
                                             0, end_bit, st));
    if (dk.Current() != w.keys_a)
        CUDA_TRY(cudaMemcpyAsync(w.keys_a, dk.Current(), 8 * (size_t)n_entries,
                                 cudaMemcpyDeviceToDevice, st));
    if (dv.Current() != w.ids_a)
        CUDA_TRY(cudaMemcpyAsync(w.ids_a, dv.Current(), 4 * (size_t)n_entries,
                                 cudaMemcpyDeviceToDevice, st));
    k_flag_heads<<<grid_for(n_entries, B), B, 0, st>>>(w.keys_a, n_entries, w.ids_b);
    CUDA_TRY(cudaGetLastError());
    tmp = w.cub_bytes;
    CUDA_TRY(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, w.ids_b, w.scan,
                                           (int)n_entries, st));
    int last = 0;
    CUDA_TRY(cudaMemcpyAsync(&last, w.scan + (n_entries - 1), sizeof(int),
                             cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *h_nnz = last;
    return 0;
}

int spfem_pattern_fill(const void *d_ws, long long n_entries,
                       long long nrows, long long ncols, long long nnz,
                       int index_bytes, void *d_indptr, void *d_indices,
                       int *d_cptr, int *d_contrib, int *d_rowstart,
                       void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (index_bytes != 4 && index_bytes != 8)
        return fail("spfem_pattern_fill: index_bytes must be 4 or 8");
    const int B = 256;
    if (d_rowstart) {
        k_fill_const<int><<<grid_for(nrows + 1, B), B, 0, st>>>(d_rowstart, nrows + 1, (int)n_entries);
        CUDA_TRY(cudaGetLastError());
    }
    if (n_entries == 0) {
        if (!d_indptr) return 0;
        if (index_bytes == 4)
            k_fill_const<int><<<grid_for(nrows + 1, B), B, 0, st>>>((int *)d_indptr, nrows + 1, 0);
        else
            k_fill_const<long long><<<grid_for(nrows + 1, B), B, 0, st>>>((long long *)d_indptr, nrows + 1, 0LL);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    PatternWs w = layout((void *)d_ws, n_entries);
    if (index_bytes == 4) {
        if (d_indptr)
            k_fill_const<int><<<grid_for(nrows + 1, B), B, 0, st>>>((int *)d_indptr, nrows + 1, (int)nnz);
        k_fill_csr<int><<<grid_for(n_entries, B), B, 0, st>>>(
            w.keys_a, w.scan, n_entries, (unsigned long long)ncols, nnz,
            (int *)d_indptr, (int *)d_indices, d_cptr, d_rowstart);
    } else {
        if (d_indptr)
            k_fill_const<long long><<<grid_for(nrows + 1, B), B, 0, st>>>((long long *)d_indptr, nrows + 1, nnz);
        k_fill_csr<long long><<<grid_for(n_entries, B), B, 0, st>>>(
            w.keys_a, w.scan, n_entries, (unsigned long long)ncols, nnz,
            (long long *)d_indptr, (long long *)d_indices, d_cptr, d_rowstart);
    }
    CUDA_TRY(cudaGetLastError());
    if (d_contrib)
        CUDA_TRY(cudaMemcpyAsync(d_contrib, w.ids_a, 4 * (size_t)n_entries,
                                 cudaMemcpyDeviceToDevice, st));
    return 0;
}

int spfem_reduce(const double *d_local, const int *d_cptr, const int *d_contrib,
                 long long nslots, double *d_values, void *stream)
{
    if (nslots <= 0) return 0;
    const int B = 256;
    k_reduce<<<grid_for(nslots, B), B, 0, (cudaStream_t)stream>>>(d_local, d_cptr, d_contrib,
                                                              nslots, d_values);
    CUDA_TRY(cudaGetLastError());
    return 0;
}


size_t spfem_morton_ws_bytes(long long nt)
{
    if (nt <= 0) nt = 1;
    size_t b1 = 0;
    cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
    cub::DoubleBuffer<int> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b1, dk, dv, (int)nt, 0, 64);
    return 2 * align_up(8 * (size_t)nt) + 2 * align_up(4 * (size_t)nt) + align_up(b1);
}

int spfem_morton_order(const double *d_p, long long ldp, int dim,
                       const int *d_t, long long ldt, int nve, long long nt,
                       const double *h_lo, const double *h_hi,
                       int *d_perm, void *d_ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (nt <= 0) return 0;
    if (dim != 2 && dim != 3) return fail("spfem_morton_order: dim must be 2 or 3");
    if (nt > 2147483647LL) return fail("spfem_morton_order: nt too large");
    if (ws_bytes < spfem_morton_ws_bytes(nt)) return fail("spfem_morton_order: workspace too small");
    char *base = (char *)d_ws;
    size_t off = 0;
    unsigned long long *ka = (unsigned long long *)(base + off); off += align_up(8 * (size_t)nt);
    unsigned long long *kb = (unsigned long long *)(base + off); off += align_up(8 * (size_t)nt);
    int *ia = (int *)(base + off); off += align_up(4 * (size_t)nt);
    int *ib = (int *)(base + off); off += align_up(4 * (size_t)nt);
    void *tmp = base + off;
    size_t tmp_bytes = ws_bytes - off;
    Box box;
    for (int d = 0; d < 3; ++d) {
        box.lo[d] = d < dim ? h_lo[d] : 0.0;
        double ext = d < dim ? (h_hi[d] - h_lo[d]) : 1.0;
        box.inv[d] = ext > 0.0 ? 1.0 / ext : 0.0;
    }
    const int B = 256;
    k_morton_keys<<<grid_for(nt, B), B, 0, st>>>(d_p, ldp, dim, d_t, ldt, nve, nt, box, ka, ia);
    CUDA_TRY(cudaGetLastError());
    cub::DoubleBuffer<unsigned long long> dk(ka, kb);
    cub::DoubleBuffer<int> dv(ia, ib);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)nt, 0,
                                             dim == 2 ? 64 : 63, st));
    CUDA_TRY(cudaMemcpyAsync(d_perm, dv.Current(), 4 * (size_t)nt,
                             cudaMemcpyDeviceToDevice, st));
    return 0;
}

} // extern "C"
