# -*- coding: utf-8 -*-
"""ctypes binding of the C-ABI library ``csrc/libspfem_b200.so``
(declared in ``include/spfem_b200.h``).

The library is built in-tree by :func:`build` (``nvcc`` for sm_100a) and is the
only way the package touches the GPU besides PyTorch tensor ownership.  A
missing library is an error -- there is no CPU fallback."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libspfem_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "--extended-lambda", "-shared", "-Xcompiler", "-fPIC",
              "-cudart", "static"]
SOURCES = ["spfem_b200.cu", "spfem_plan.cu"]


class SpfemArgs(C.Structure):
    """Mirror of ``struct SpfemArgs`` (include/spfem_b200.h)."""
    _fields_ = [
        ("n", C.c_longlong), ("ldo", C.c_longlong),
        ("sel", C.c_void_p),
        ("t", C.c_void_p), ("ldt", C.c_longlong),
        ("p", C.c_void_p), ("ldp", C.c_longlong),
        ("tdof_u", C.c_void_p), ("ldd", C.c_longlong),
        ("interp", C.c_void_p * 8),
        ("fv", C.c_void_p), ("e1", C.c_void_p), ("e2", C.c_void_p),
        ("lf1", C.c_void_p),
        ("out", C.c_void_p),
        ("ldk", C.c_longlong),
        ("params", C.c_double * 32),
        ("fields", C.c_void_p * 8),
        ("ldf", C.c_longlong),
    ]

    def set_params(self, plan):
        """Values of the form's run-time scalars (``codegen.Emitter.params``)."""
        for i, v in enumerate(getattr(plan.spec, "params", ())):
            self.params[i] = v


class SpfemTileArgs(C.Structure):
    """Mirror of ``struct SpfemTileArgs`` (include/spfem_b200.h; the generated
    source declares the identical struct, codegen.TILE_KERNEL)."""
    _fields_ = [
        ("base", SpfemArgs),
        ("desc", C.c_void_p), ("blob", C.c_void_p), ("values", C.c_void_p),
        ("ldl", C.c_int), ("max_a", C.c_int), ("max_b", C.c_int),
        ("compact", C.c_int), ("store", C.c_int), ("prefetch", C.c_int),
        ("sorted", C.c_int), ("els", C.c_int),
    ]


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
PLAN_MAX_BATCHES = 256


class SpfemPlanIn(C.Structure):
    """Mirror of ``struct SpfemPlanIn`` (include/spfem_b200.h)."""
    _fields_ = [
        ("rowdof", C.c_void_p), ("ldr", C.c_longlong), ("nloc_v", C.c_int),
        ("coldof", C.c_void_p), ("ldc", C.c_longlong), ("nloc_u", C.c_int),
        ("n", C.c_longlong), ("nrows", C.c_longlong), ("ncols", C.c_longlong),
        ("vt", C.c_void_p), ("ldvt", C.c_longlong), ("nve", C.c_int), ("dim", C.c_int),
        ("bsv", C.c_int), ("bsu", C.c_int),
        ("alias", C.c_void_p), ("nuniq", C.c_int),
        ("E", C.c_int), ("with_el", C.c_int), ("sort_slots", C.c_int), ("compact", C.c_int),
        ("row_mask", C.c_void_p), ("index_bytes", C.c_int), ("batch_cap", C.c_longlong),
        ("split_factor", C.c_double), ("ne_budget", C.c_int), ("owner_rule", C.c_int),
    ]


class SpfemPlanBatch(C.Structure):
    _fields_ = [
        ("desc", C.c_void_p), ("blob", C.c_void_p), ("vid", C.c_void_p), ("vptr", C.c_void_p),
        ("ntiles", C.c_longlong), ("blob_bytes", C.c_longlong), ("nvid", C.c_longlong),
        ("first_tile", C.c_longlong), ("nslots", C.c_longlong),
    ]


class SpfemPlanOut(C.Structure):
    _fields_ = [
        ("nnz", C.c_longlong), ("indptr", C.c_void_p), ("indices", C.c_void_p),
        ("nbatches", C.c_int), ("ldl", C.c_int), ("els", C.c_int),
        ("bsv", C.c_int), ("bsu", C.c_int),
        ("max_ne", C.c_int), ("max_nr", C.c_int), ("max_ns", C.c_int), ("max_store", C.c_int),
        ("max_bytes_a", C.c_int), ("max_bytes_b", C.c_int),
        ("ntiles", C.c_longlong), ("n_pairs", C.c_longlong), ("n_contrib", C.c_longlong),
        ("batches", SpfemPlanBatch * PLAN_MAX_BATCHES),
    ]


class SpfemFanPlanIn(C.Structure):
    """Mirror of ``struct SpfemFanPlanIn`` (include/spfem_b200.h)."""
    _fields_ = [
        ("p", C.c_void_p), ("ldp", C.c_longlong), ("nv", C.c_longlong),
        ("t", C.c_void_p), ("ldt", C.c_longlong), ("nt", C.c_longlong),
        ("indptr", C.c_void_p), ("indices", C.c_void_p), ("index_bytes", C.c_int),
        ("rows_per_tile", C.c_int), ("tiny", C.c_int), ("row_order", C.c_int),
        ("lo", C.c_double * 2), ("hi", C.c_double * 2),
    ]


class SpfemFanPlanOut(C.Structure):
    _fields_ = [
        ("desc", C.c_void_p), ("blob", C.c_void_p), ("blob_bytes", C.c_longlong),
        ("vid", C.c_void_p), ("pc_index", C.c_void_p),
        ("nvid", C.c_longlong), ("ntiles", C.c_longlong),
        ("max_a", C.c_int), ("max_ns", C.c_int), ("max_neighbours", C.c_int), ("format", C.c_int),
        ("halo_factor", C.c_double),
    ]


class SpfemFanArgs(C.Structure):
    """Mirror of ``struct SpfemFanArgs`` (generated source, codegen.FAN_KERNEL)."""
    _fields_ = [
        ("base", SpfemArgs),
        ("desc", C.c_void_p), ("blob", C.c_void_p), ("values", C.c_void_p),
        ("max_a", C.c_int), ("max_b", C.c_int), ("max_ns", C.c_int), ("prefetch", C.c_int),
    ]


SYMBOLS = {
    # name: (restype, argtypes)
    "spfem_abi_version": (C.c_int, []),
    "spfem_last_error": (C.c_char_p, []),
    "spfem_device_count": (C.c_int, []),
    "spfem_nvrtc_compile": (C.c_int, [C.c_char_p, C.c_char_p,
                                      C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_void_p)]),
    "spfem_free": (None, [C.c_void_p]),
    "spfem_module_load": (C.c_int, [C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_void_p)]),
    "spfem_module_unload": (C.c_int, [C.c_void_p]),
    "spfem_module_get_kernel": (C.c_int, [C.c_void_p, C.c_char_p,
                                          C.POINTER(C.c_void_p)]),
    "spfem_launch": (C.c_int, [C.c_void_p, C.POINTER(SpfemArgs), C.c_int,
                               C.c_void_p]),
    "spfem_launch_tiled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint,
                                     C.c_size_t, C.c_void_p]),
    "spfem_pattern_ws_bytes": (C.c_size_t, [C.c_longlong]),
    "spfem_pattern_sort": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p,
                                     C.c_longlong, C.c_longlong, C.c_int,
                                     C.c_int, C.c_longlong, C.c_longlong,
                                     C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_longlong), C.c_void_p]),
    "spfem_pattern_fill": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong,
                                     C.c_longlong, C.c_longlong, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "spfem_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_longlong, C.c_void_p, C.c_void_p]),
    "spfem_tileplan_build": (C.c_int, [C.POINTER(SpfemPlanIn), C.POINTER(SpfemPlanOut),
                                       ALLOC_FN, FREE_FN, C.c_void_p, C.c_int, C.c_void_p]),
    "spfem_fanplan_build": (C.c_int, [C.POINTER(SpfemFanPlanIn), C.POINTER(SpfemFanPlanOut),
                                      ALLOC_FN, FREE_FN, C.c_void_p, C.c_int, C.c_void_p]),
    "spfem_tileplan_fill_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                             C.c_int, C.c_void_p]),
    "spfem_morton_ws_bytes": (C.c_size_t, [C.c_longlong]),
    "spfem_morton_order": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int,
                                     C.c_void_p, C.c_longlong, C.c_int,
                                     C.c_longlong, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


class SpfemError(RuntimeError):
    """An entry point of libspfem_b200 returned a non-zero status."""


def build(force=False, verbose=False):
    """Compile ``csrc/spfem_b200.cu`` into ``csrc/libspfem_b200.so``."""
    srcs = [os.path.join(CSRC, f) for f in SOURCES]
    hdr = os.path.join(INCLUDE, "spfem_b200.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(f) for f in srcs + [hdr])):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + srcs + ["-lnvrtc", "-Xlinker",
                                 "-rpath", "-Xlinker", "/usr/local/cuda/lib64"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


def load():
    """Load the library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpfemError("libspfem_b200.so is not built: run "
                         "`python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.spfem_abi_version() != 4:
        raise SpfemError("libspfem_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise SpfemError(load().spfem_last_error().decode("utf-8", "replace"))


def nvrtc_compile(source, arch="sm_100a"):
    """CUDA C -> cubin bytes through ``spfem_nvrtc_compile``."""
    lib = load()
    image = C.c_void_p()
    size = C.c_size_t()
    log = C.c_void_p()
    st = lib.spfem_nvrtc_compile(source.encode("utf-8"), arch.encode("ascii"),
                                 C.byref(image), C.byref(size), C.byref(log))
    try:
        if st != 0:
            raise SpfemError(lib.spfem_last_error().decode("utf-8", "replace"))
        return C.string_at(image.value, size.value)
    finally:
        if image.value:
            lib.spfem_free(image)
        if log.value:
            lib.spfem_free(log)
