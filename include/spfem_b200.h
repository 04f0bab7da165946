/* spfem_b200 -- C ABI of the B200-native element-assembly path.
 *
 * The reference (kinnala/sp.fem) is pure Python: it has no FFI.  This header
 * is the boundary a maintainer would bind from `spfem/assembly.py` (ctypes
 * stub in INTEGRATION.md) to replace the bodies of
 *
 *   AssemblerElement.iasm   spfem/assembly.py:919-1006
 *   AssemblerElement.fasm   spfem/assembly.py:1011-1183
 *   coo_matrix(...).tocsr() spfem/assembly.py:985-986, 1158-1159 (SciPy)
 *   coo_matrix(...).toarray().T[0]         :1005-1006, 1182-1183
 *
 * Conventions: every function returns 0 on success and a non-zero code on
 * failure; `spfem_last_error()` then returns a thread-local message.  All
 * `d_*` pointers are DEVICE pointers owned by the caller (PyTorch tensors in
 * the Python host layer); `h_*` are host pointers.  `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = legacy default stream).  No entry
 * point allocates device memory behind the caller's back except the `*_host`
 * convenience call, which says so.  Re-entrant per stream.
 */
#ifndef SPFEM_B200_H
#define SPFEM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPFEM_ABI_VERSION 4

/* Argument block of every generated kernel (see spfem_b200/codegen.py; the
 * generated source declares the identical struct). */
typedef struct SpfemArgs {
    long long n;             /* elements / facets to process                    */
    long long ldo;           /* leading dimension of the entry-major output      */
    const int *sel;          /* optional element subset (`tind`) or NULL          */
    const int *t;            /* vertex connectivity t[k*ldt + e]                  */
    long long ldt;
    const double *p;         /* vertex coordinates p[k*ldp + v]                   */
    long long ldp;
    const int *tdof_u;       /* trial DOF table (interp only) [j*ldd + e]         */
    long long ldd;
    const double *interp[8]; /* interpolated solution vectors (`w`)               */
    const int *fv;           /* facet kernels: facet vertex ids fv[k*n + f]       */
    const int *e1;           /* facet kernels: first / second adjacent element    */
    const int *e2;
    const int *lf1;          /* facet kernels: local facet index inside e1        */
    double *out;             /* local entries out[entry*ldo + k*ldk]              */
    long long ldk;           /* stride between items inside an entry plane (1)     */
    double params[32];       /* run-time scalars of the traced form (captured Python
                                numbers that are not structural constants: a time
                                step, a penalty, ...), in the order codegen assigned  */
    const double *fields[8]; /* host-evaluated coefficient fields (opaque callables of
                                x / h / w, spfem/mesh.py:885-901): fields[i][q*ldf + e],
                                q = quadrature point, e = element in device order      */
    long long ldf;
} SpfemArgs;

/* Argument block of the fused tile kernel `<name>_tile` (codegen.TILE_KERNEL);
 * one launch per SpfemPlanBatch of the tile plan (see below). */
typedef struct SpfemTileArgs {
    SpfemArgs base;
    const int *desc;             /* SpfemPlanBatch.desc                              */
    const unsigned char *blob;   /* SpfemPlanBatch.blob                              */
    double *values;              /* CSR values                                        */
    int ldl;                     /* SpfemPlanOut.ldl                                  */
    int max_a, max_b;            /* SpfemPlanOut.max_bytes_a / max_bytes_b            */
    int compact;                 /* SpfemPlanIn.compact                               */
    int store;                   /* doubles of shared memory for the local values:
                                    nuniq * ldl, or max_store + 2 (compact)           */
    int prefetch;                /* L2 prefetch distance in tiles (0 = off)           */
    int sorted;                  /* SpfemPlanIn.sort_slots != 0: every slot carries its CSR
                                    position; 0: slots in CSR order, positions from the
                                    row records + 32-slot headers of section B         */
    int els;                     /* SpfemPlanOut.els                                  */
} SpfemTileArgs;

/* Argument block of the vertex-fan kernels `<name>_fan*` (codegen.FAN_KERNEL). */
typedef struct SpfemFanArgs {
    SpfemArgs base;
    const int *desc;             /* 8 ints per tile: offset/16, bytes A, bytes B, rows, slots, vertices */
    const unsigned char *blob;
    double *values;
    int max_a, max_b, max_ns, prefetch;
} SpfemFanArgs;

int         spfem_abi_version(void);
const char *spfem_last_error(void);
/* number of visible CUDA devices (0 without a driver); never fails */
int         spfem_device_count(void);

/* ---- run-time compilation (NVRTC, --gpu-architecture=sm_100a) -------------
 * Compiles `src` to a cubin.  `*image` is malloc'ed (release with spfem_free),
 * `*log` likewise (may be NULL on success).  Works without a GPU.            */
int  spfem_nvrtc_compile(const char *src, const char *arch, char **image,
                         size_t *image_size, char **log);
void spfem_free(void *ptr);

/* ---- module handling / launch ------------------------------------------- */
int spfem_module_load(const void *image, size_t image_size, void **module);
int spfem_module_unload(void *module);
int spfem_module_get_kernel(void *module, const char *name, void **kernel);
/* Launches a generated kernel: one thread per item, `block` threads per CTA. */
int spfem_launch(void *kernel, const SpfemArgs *args, int block, void *stream);
/* Launches the fused tile kernel `<name>_tile` of a generated module: `args`
 * points to a SpfemTileArgs / SpfemFanArgs block, one CTA per tile, `smem_bytes` of dynamic shared memory
 * (opt-in above 48 KB is handled here).                                       */
int spfem_launch_tiled(void *kernel, const void *args, unsigned grid, unsigned block,
                       size_t smem_bytes, void *stream);

/* ---- sparsity pattern: COO (row, col) pairs -> CSR + contribution lists ----
 * Replaces SciPy's coo_tocsr + sum_duplicates structure pass.  Entry ids are
 * `id = (j*nloc_v + i)*n + k` for item k, local trial j, local test i
 * (the reference's COO block order, spfem/assembly.py:977); entry `id` has
 * row = d_rowdof[i*ldr + k], col = d_coldof[j*ldc + k] (col = 0 if
 * d_coldof == NULL: linear forms).
 *
 * spfem_pattern_sort : sorts all entries by (row, col) [stable radix sort on
 *    64-bit keys], flags the first entry of every distinct (row, col) and
 *    returns their number in *h_nnz (synchronises `stream`).
 * spfem_pattern_fill : writes indptr (nrows+1), indices (nnz), cptr (nnz+1)
 *    and contrib (n_entries): CSR slot s sums the local entries
 *    contrib[cptr[s] .. cptr[s+1]) in that (ascending entry id) order.
 *    Optional d_rowstart (nrows+1): position in `contrib` where the entries
 *    of each row start -- with it spfem_reduce(cptr = rowstart, nslots =
 *    nrows) yields the dense load vector of a linear form (toarray()).
 *    d_indptr / d_indices / d_cptr may be NULL when only rowstart is wanted.
 * `index_bytes` is 4 or 8 (SciPy's rule: int64 iff max(n_entries, nrows,
 * ncols) > 2^31-1).                                                           */
size_t spfem_pattern_ws_bytes(long long n_entries);
int spfem_pattern_sort(const int *d_rowdof, long long ldr,
                       const int *d_coldof, long long ldc,
                       long long n, int nloc_v, int nloc_u,
                       long long nrows, long long ncols,
                       void *d_ws, size_t ws_bytes, long long *h_nnz,
                       void *stream);
int spfem_pattern_fill(const void *d_ws, long long n_entries,
                       long long nrows, long long ncols, long long nnz,
                       int index_bytes, void *d_indptr, void *d_indices,
                       int *d_cptr, int *d_contrib, int *d_rowstart,
                       void *stream);

/* ---- deterministic duplicate summation ------------------------------------
 * d_values[s] = sum_{k = cptr[s]}^{cptr[s+1]-1} d_local[contrib[k]], added in
 * ascending k with plain fp64 adds (fixed order => bitwise reproducible).    */
int spfem_reduce(const double *d_local, const int *d_cptr,
                 const int *d_contrib, long long nslots, double *d_values,
                 void *stream);

/* ---- element ordering -------------------------------------------------------
 * Morton (Z-order) permutation of the elements by centroid: d_perm[k] = index
 * of the k-th element along the curve.  ws: spfem_morton_ws_bytes(nt).        */
size_t spfem_morton_ws_bytes(long long nt);
int spfem_morton_order(const double *d_p, long long ldp, int dim,
                       const int *d_t, long long ldt, int nve, long long nt,
                       const double *h_lo, const double *h_hi, /* bounding box */
                       int *d_perm, void *d_ws, size_t ws_bytes, void *stream);


/* ---- tile plan: sparsity pattern + plan of the fused kernel, natively -------
 * (csrc/spfem_plan.cu).  Builds, for the element range described by the DOF
 * tables, (i) the CSR structure of SciPy's coo_matrix(...).tocsr()
 * (spfem/assembly.py:985-986: sorted unique columns per row, explicit zeros
 * kept) and (ii) the per-tile blobs the generated kernel `<name>_tile` reads
 * (codegen.TILE_KERNEL): tiles of E consecutive elements, every row owned by
 * the tile holding most of its elements (SpfemPlanIn.owner_rule), a tile's
 * element list = its elements + the halo elements touching its rows.  The work is done in batches of tiles with at
 * most `batch_cap` contributions each, so the memory in use is bounded by the
 * batch, not by the mesh (no 2^31 limit on Nu*Nv*nt).  One SpfemPlanBatch = one
 * launch of the tile kernel.
 *
 * Vector elements: with bsv / bsu > 1 the DOF tables are checked for the
 * interleaved structure dof = bs * scalar_dof + component; if it holds, the
 * plan is built on the scalar structure and one slot stands for a bsv x bsu
 * block of CSR values (SpfemPlanOut.bsv / bsu report what was used).
 *
 * Memory: every buffer is requested from the caller through `alloc` (tag says
 * what for; tags starting with "out:" are results whose ownership passes to
 * the caller, everything else is released through `release` before return).
 * on_device = 0: the same code on host pointers (tests without a GPU).
 * Return codes: 0 ok, 2 = a tile does not fit the 16-bit fields (retry with a
 * smaller E), other = error (spfem_last_error()).                              */
typedef void *(*spfem_alloc_fn)(void *ctx, size_t bytes, const char *tag);
typedef void (*spfem_free_fn)(void *ctx, void *ptr);

#define SPFEM_PLAN_MAX_BATCHES 256

typedef struct SpfemPlanIn {
    const int *rowdof;        /* test DOF table  rowdof[i*ldr + e], device element order */
    long long ldr;
    int nloc_v;
    const int *coldof;        /* trial DOF table (NULL: linear form, one column)      */
    long long ldc;
    int nloc_u;
    long long n;              /* elements                                              */
    long long nrows, ncols;
    const int *vt;            /* vertex table vt[r*ldvt + e]                           */
    long long ldvt;
    int nve, dim;
    int bsv, bsu;             /* block sizes to try (components of a vector element)   */
    const int *alias;         /* [nloc_u*nloc_v] local entry j*nloc_v+i -> distinct    */
    int nuniq;                /*   local value (scalar plans only), their number       */
    int E;                    /* elements per tile                                     */
    int with_el;              /* keep element ids in the blobs (interp gathers)        */
    int sort_slots;           /* 0: CSR order (section B: row records + 32-slot headers),
                                 1: by contribution count, 2: coarse classes (section B:
                                 per-slot positions); layout in spfem_b200/tileplan.py */
    int compact;              /* shared memory holds only the local values that are read */
    const unsigned char *row_mask; /* optional [nrows]: 0 = row is not assembled here  */
    int index_bytes;          /* 4 or 8: width of indptr / indices                     */
    long long batch_cap;      /* contributions per batch (0: 2^27)                     */
    double split_factor;      /* > 0: tiles whose element list (with halo) exceeds
                                 split_factor x the median are split (the Morton curve
                                 jumps; the largest tile sizes every CTA's shared memory) */
    int ne_budget;            /* > 0: split tiles with more than ne_budget elements
                                 (absolute form of the same rule; never below the median) */
    int owner_rule;           /* owner of a CSR row: 0 the lowest tile touching it, 1 the
                                 tile holding most of its elements (ties: the lowest)     */
} SpfemPlanIn;

typedef struct SpfemPlanBatch {
    void *desc;               /* int32 [16 * ntiles]                                    */
    void *blob;               /* bytes, 16-byte aligned sections                        */
    void *vid;                /* int32 [nvid] global vertex id of every tile vertex     */
    void *vptr;               /* int64 [ntiles + 1] first tile vertex of every tile     */
    long long ntiles, blob_bytes, nvid, first_tile, nslots;
} SpfemPlanBatch;

typedef struct SpfemPlanOut {
    long long nnz;
    void *indptr, *indices;   /* index_bytes wide; indices NULL for a linear form       */
    int nbatches;
    int ldl;                  /* leading dimension of the shared-memory value image     */
    int els;                  /* block plans: contribution code = (entry << els) | elem */
    int bsv, bsu;
    int max_ne, max_nr, max_ns, max_store, max_bytes_a, max_bytes_b;
    long long ntiles, n_pairs, n_contrib;
    SpfemPlanBatch batches[SPFEM_PLAN_MAX_BATCHES];
} SpfemPlanOut;

int spfem_tileplan_build(const SpfemPlanIn *in, SpfemPlanOut *out, spfem_alloc_fn alloc,
                         spfem_free_fn release, void *alloc_ctx, int on_device, void *stream);
/* (Re)writes the tiles' private vertex coordinates from p[k*ldp + v].            */
int spfem_tileplan_fill_coords(const int *desc, unsigned char *blob, const int *vid,
                               const long long *vptr, long long ntiles, const double *p,
                               long long ldp, int dim, int on_device, void *stream);

/* ---- vertex-fan plan: plan of `<name>_fan*` (codegen.FAN_KERNEL), natively ----------
 * Scalar P1 on triangles (one DOF per vertex; CSR rows = vertex stars).  For
 * every vertex the counter-clockwise ring of its neighbours is walked from the
 * triangles' directed edges (no facet tables needed); vertices are cut into
 * tiles of `rows_per_tile` along a Morton curve, the rows of a tile are in CSR
 * order, and every tile gets one blob: interleaved (x, y) of its vertices (rows
 * first, then the halo vertices in id order -- filled through vid / pc_index),
 * then the row records (spfem_b200/fanplan.py documents the three formats).
 * Memory through the allocator callbacks as above.  Return code 5: the mesh /
 * pattern is not fan shaped (non-manifold, valence > 12, ...): use the tile plan. */
typedef struct SpfemFanPlanIn {
    const double *p;          /* coordinates p[k*ldp + v], k = 0, 1                    */
    long long ldp, nv;
    const int *t;             /* connectivity t[k*ldt + e], k = 0..2 (any element order) */
    long long ldt, nt;
    const void *indptr;       /* CSR pattern of the vertex x vertex matrix              */
    const void *indices;
    int index_bytes;          /* 4 or 8                                                  */
    int rows_per_tile;
    int tiny;                 /* allow the 20-byte records (<= 256 vertices per tile)    */
    int row_order;            /* 0: rows of a tile in CSR order, 1: in Morton order      */
    double lo[2], hi[2];      /* bounding box of the vertices                            */
} SpfemFanPlanIn;

typedef struct SpfemFanPlanOut {
    void *desc;               /* int32 [8 * ntiles]: offset/16, bytes A, bytes B, rows,
                                 slots, vertices, 0, 0                                   */
    void *blob;
    long long blob_bytes;
    void *vid;                /* int32 [nvid] vertex of every coordinate slot            */
    void *pc_index;           /* int64 [nvid] index (in doubles) of its x in the blob    */
    long long nvid, ntiles;
    int max_a;                /* largest blob (bytes)                                    */
    int max_ns;               /* most CSR slots of one tile                              */
    int max_neighbours;
    int format;               /* 0: 20-byte records, 1: 32-byte, 2: 48-byte              */
    double halo_factor;
} SpfemFanPlanOut;

int spfem_fanplan_build(const SpfemFanPlanIn *in, SpfemFanPlanOut *out, spfem_alloc_fn alloc,
                        spfem_free_fn release, void *alloc_ctx, int on_device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPFEM_B200_H */
