"""ctypes binding of libssdr_b200.so (the C ABI declared in include/ssdr_b200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssdr_b200.so")
_lib = None

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)
sz = C.c_size_t
vp = C.c_void_p


class KnnStats(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("tie_rows", C.c_uint64), ("tree_builds", C.c_uint64),
                ("dist_evals", C.c_uint64), ("grid_build_ms", C.c_double), ("main_kernel_ms", C.c_double),
                ("tie_path_ms", C.c_double), ("tree_build_ms", C.c_double), ("kernel_launches", C.c_uint64)]


# name -> argtypes (every function returns int status unless listed in _RESTYPE)
SIGNATURES = {
    "ssdr_last_error": [],
    "ssdr_version": [],
    "ssdr_device_count": [C.POINTER(C.c_int)],
    "ssdr_set_device": [C.c_int],
    "ssdr_get_device": [C.POINTER(C.c_int)],
    "ssdr_device_sm_count": [C.POINTER(C.c_int)],
    "ssdr_synchronize": [],
    "ssdr_host_alloc": [C.POINTER(vp), sz],
    "ssdr_host_free": [vp],
    "ssdr_knn": [vp, sz, sz, vp, sz, sz, vp],
    "ssdr_knn_batch": [vp, sz, sz, sz, vp, sz, sz, vp],
    "ssdr_knn_batch_strided": [vp, sz, sz, sz, sz, vp, sz, sz, sz, vp],
    "ssdr_knn_batch_dev": [vp, sz, sz, vp, sz, sz, vp, vp, C.POINTER(KnnStats)],
    "ssdr_knn_batch_dev_i32": [vp, sz, sz, vp, sz, sz, vp, vp, C.POINTER(KnnStats)],
    "ssdr_knn_pyramid_dev": [vp, sz, sz, vp, sz, sz, vp, vp, vp],
    "ssdr_knn_pyramid": [vp, sz, sz, sz, vp, sz, sz, vp, vp],
    "ssdr_knn_status": [vp],
    "ssdr_gcn_adjacency_f64": [sz, sz, vp, vp, vp, vp, C.POINTER(vp)],
    "ssdr_gcn_fetch": [vp, vp],
    "ssdr_gcn_free": [vp],
    "ssdr_gcn_propagate_f64": [vp, vp, sz, vp, sz, C.c_int, C.c_int, vp],
    "ssdr_knn_pyramid_launches": [],
    "ssdr_knn_debug_tree": [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "ssdr_knn_debug_build_timing": [vp, sz, sz, vp],
    "ssdr_grid_subsample": [vp, vp, vp, sz, sz, sz, C.c_float, C.c_int, C.POINTER(sz), C.POINTER(vp)],
    "ssdr_grid_fetch": [vp, vp, vp, vp],
    "ssdr_grid_fetch_ex": [vp, vp, vp, vp, vp, vp],
    "ssdr_grid_free": [vp],
    "ssdr_grid_debug_timing": [vp, C.POINTER(C.c_int)],
    "ssdr_grid_subsample_typed": [vp, vp, C.c_int, vp, C.c_int, sz, sz, sz, C.c_float, C.c_int, C.POINTER(sz),
                                  C.POINTER(vp)],
    "ssdr_grid_subsample_dev": [vp, vp, vp, sz, sz, sz, C.c_float, C.c_int, vp, C.POINTER(sz), C.POINTER(vp)],
    "ssdr_grid_dev_ptrs": [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
    "ssdr_grid_subsample_slab_dev": [vp, vp, vp, sz, sz, sz, C.c_float, C.c_int, vp, C.c_int, C.c_ulonglong,
                                     C.c_ulonglong, vp, C.POINTER(sz), C.POINTER(vp)],
    "ssdr_grid_subsample_into_dev": [vp, vp, vp, sz, sz, sz, C.c_float, vp, C.c_int, C.c_ulonglong, C.c_ulonglong, vp, vp,
                                     vp, vp, vp, sz, vp, C.POINTER(sz)],
    "ssdr_grid_bbox_dev": [vp, sz, vp, vp],
    "ssdr_grid_point_layers_dev": [vp, sz, vp, C.c_float, C.c_int, vp, vp, C.POINTER(C.c_ulonglong)],
    "ssdr_grid_layer_hist_dev": [vp, sz, vp, C.c_float, C.c_int, vp, sz, sz, vp],
    "ssdr_grid_route_dev": [vp, vp, vp, sz, sz, sz, C.c_float, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp],
    "ssdr_fps_f32": [vp, sz, sz, C.c_int32, sz, vp],
    "ssdr_fps_f64": [vp, sz, sz, C.c_int32, sz, vp],
    "ssdr_fps_f32_dev": [vp, sz, sz, C.c_int32, sz, vp, vp],
    "ssdr_fps_f64_dev": [vp, sz, sz, C.c_int32, sz, vp, vp],
    "ssdr_kcenter_f32": [vp, sz, sz, vp, sz, sz, vp],
    "ssdr_kcenter_f64": [vp, sz, sz, vp, sz, sz, vp],
    "ssdr_kcenter_f32_dev": [vp, sz, sz, vp, sz, sz, vp, vp],
    "ssdr_kcenter_f64_dev": [vp, sz, sz, vp, sz, sz, vp, vp],
    "ssdr_chamfer_matrix_f64": [vp, vp, sz, vp],
    "ssdr_superpoint_fps_f64": [vp, vp, sz, vp, C.c_int32, sz, vp],
    "ssdr_superpoint_fps": [vp, vp, sz, vp, C.c_int, C.c_int32, sz, vp],
    "ssdr_chamfer_matrix_f64_dev": [vp, vp, vp, sz, vp, vp],
    "ssdr_fps_f32_sharded": [vp, sz, sz, sz, sz, C.c_int32, sz, vp, vp, vp],
    "ssdr_peer_group_create": [C.c_int, C.c_int, C.POINTER(vp)],
    "ssdr_peer_group_export": [vp, vp],
    "ssdr_peer_group_connect": [vp, vp],
    "ssdr_peer_group_connect_local": [C.POINTER(vp), C.c_int],
    "ssdr_peer_group_destroy": [vp],
    "ssdr_peer_group_check": [vp, vp],
    "ssdr_fps_sharded_p2p": [C.c_int, vp, sz, sz, sz, sz, C.c_int32, sz, vp, vp, vp, C.c_int],
    "ssdr_kcenter_sharded_p2p": [C.c_int, vp, sz, sz, sz, sz, vp, sz, sz, vp, vp, vp, C.c_int],
    "ssdr_nccl_unique_id": [vp],
    "ssdr_nccl_comm_init": [C.POINTER(vp), C.c_int, vp, C.c_int],
    "ssdr_nccl_comm_destroy": [vp],
}
_RESTYPE = {"ssdr_last_error": C.c_char_p, "ssdr_knn_pyramid_launches": C.c_ulonglong}


def lib():
    """Load the shared library; fail loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libssdr_b200.so is not built (%s). Run `python -m ssdr_al_b200.build` -- this package has no "
                "CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = L
    return _lib


def check(status):
    if status != 0:
        msg = lib().ssdr_last_error()
        raise RuntimeError((msg or b"libssdr_b200 error").decode("utf-8", "replace"))


def ptr(a):
    """Address of a numpy array for a c_void_p argument (None -> NULL).  The integer from __array_interface__ is
    several times cheaper than a.ctypes.data_as(...), which matters for the small calls of the pyramid's lower levels;
    the caller keeps `a` alive across the foreign call."""
    return None if a is None else a.__array_interface__["data"][0]


def device_count():
    n = C.c_int(0)
    lib().ssdr_device_count(C.byref(n))
    return n.value


def set_device(i):
    check(lib().ssdr_set_device(int(i)))


class _PinnedBlock(object):
    """A CUDA pinned host allocation that returns itself to the pool when the last numpy view dies."""
    __slots__ = ("ptr", "nbytes", "__array_interface__", "__weakref__")

    def __init__(self, ptr, nbytes, shape, typestr):
        self.ptr = ptr
        self.nbytes = nbytes
        self.__array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

    def __del__(self):
        global _POOL_BYTES
        try:
            if _POOL_BYTES + self.nbytes > _POOL_LIMIT:  # the pool is full: give the pages back to the OS
                _lib.ssdr_host_free(vp(self.ptr))
            else:
                _POOL.setdefault(self.nbytes, []).append(self.ptr)
                _POOL_BYTES += self.nbytes
        except Exception:  # interpreter shutdown
            pass


_POOL = {}          # size class -> free pinned pointers
_POOL_BYTES = 0     # bytes parked in _POOL (blocks in use by live arrays are not counted)
_POOL_LIMIT = int(os.environ.get("SSDR_PINNED_POOL_BYTES", 1 << 30))


def pinned_pool_bytes():
    return _POOL_BYTES


def pinned_pool_trim():
    """Release every parked pinned block (long-running data preparation can call this between scans)."""
    global _POOL_BYTES
    for cls, ptrs in list(_POOL.items()):
        while ptrs:
            lib().ssdr_host_free(vp(ptrs.pop()))
    _POOL_BYTES = 0


def pinned_empty(shape, dtype):
    """Fresh numpy array in CUDA pinned host memory (device<->host copies run at full PCIe speed, no page faults).
    The memory goes back to a pool when the array is garbage collected, so every call still returns an independent
    array like np.empty does."""
    dtype = np.dtype(dtype)
    shape = tuple(int(v) for v in shape)
    nbytes = dtype.itemsize
    for v in shape:
        nbytes *= v
    if nbytes == 0:
        return np.empty(shape, dtype)
    global _POOL_BYTES
    # size classes: powers of two up to 1 MiB, then multiples of 1/8 of the next lower power of two (<= 12.5 % slack
    # instead of up to 100 % on large results)
    if nbytes <= (1 << 20):
        cls = 1 << max(12, (nbytes - 1).bit_length())
    else:
        step = 1 << ((nbytes - 1).bit_length() - 4)
        cls = (nbytes + step - 1) // step * step
    free = _POOL.get(cls)
    if free:
        ptr = free.pop()
        _POOL_BYTES -= cls
    else:
        p = vp()
        check(lib().ssdr_host_alloc(C.byref(p), cls))
        ptr = p.value
    return np.asarray(_PinnedBlock(ptr, cls, shape, dtype.str))  # one array object; its base keeps the block alive


def pinned_zeros(shape, dtype):
    a = pinned_empty(shape, dtype)
    a[...] = 0
    return a
