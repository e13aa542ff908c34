"""Grid subsampling behind the reference's `grid_subsampling.compute` interface
(utils/cpp_wrappers/cpp_subsampling/wrapper.cpp:58-286): same keyword-only arguments, conversions, error
messages, return arity, shapes and dtypes."""
import ctypes as C

import numpy as np

from . import _lib

ORDER_KEY = 0
ORDER_REFERENCE = 1
DTYPE_NATIVE = 0
DTYPE_U8 = 1
_default_order = ORDER_KEY


def set_default_order(order):
    """"key" (ascending voxel key, default) or "reference" (the reference's hash-iteration row order)."""
    global _default_order
    _default_order = {"key": ORDER_KEY, "reference": ORDER_REFERENCE}[order]


def _to_array(obj, dtype, what):
    try:
        return np.ascontiguousarray(obj, dtype=dtype)  # PyArray_FROM_OTF(..., NPY_IN_ARRAY)  wrapper.cpp:100-106
    except Exception:
        raise RuntimeError("Error converting input %s to numpy arrays of type %s" % (what, np.dtype(dtype).name))


def compute(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0, order=None,
            return_keys=False):
    # wrapper.cpp:86-90 -- validated, then ignored (always barycenters)
    if method not in ("barycenters", "voxelcenters"):
        raise RuntimeError("Error parsing method. Valid method names are \"barycenters\" and \"voxelcenters\" ")
    use_feature = features is not None
    use_classes = classes is not None
    pts = _to_array(points, np.float32, "points")
    # uint8 colours / labels (what every data-prep caller passes) travel as bytes and are widened on the device --
    # the values are the ones PyArray_FROM_OTF(NPY_FLOAT / NPY_INT) would produce (wrapper.cpp:100-106)
    f_u8 = use_feature and isinstance(features, np.ndarray) and features.dtype == np.uint8
    c_u8 = use_classes and isinstance(classes, np.ndarray) and classes.dtype == np.uint8
    feats = _to_array(features, np.uint8 if f_u8 else np.float32, "features") if use_feature else None
    cls = _to_array(classes, np.uint8 if c_u8 else np.int32, "classes") if use_classes else None
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    if use_feature and feats.ndim != 2:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if use_classes and cls.ndim > 2:
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    N = pts.shape[0]
    fdim = feats.shape[1] if use_feature else 0
    ldim = 1
    if use_classes and cls.ndim == 2:
        ldim = cls.shape[1]
    if use_feature and feats.shape[0] != N:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if use_classes and (cls.ndim == 0 or cls.shape[0] != N):
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    if verbose > 0:
        print("Computing cloud pyramid with support points: ")
    if N == 0:
        raise RuntimeError("Error")  # the reference reads points[0] (UB); an empty result is "Error" wrapper.cpp:227

    L = _lib.lib()
    M = C.c_size_t(0)
    h = C.c_void_p()
    _lib.check(L.ssdr_grid_subsample_typed(_lib.ptr(pts), _lib.ptr(feats), DTYPE_U8 if f_u8 else DTYPE_NATIVE,
                                           _lib.ptr(cls), DTYPE_U8 if c_u8 else DTYPE_NATIVE, N, fdim,
                                           ldim if use_classes else 0, float(sampleDl),
                                           _default_order if order is None else
                                           {"key": ORDER_KEY, "reference": ORDER_REFERENCE}.get(order, order),
                                           C.byref(M), C.byref(h)))
    try:
        m = M.value
        if m < 1:
            raise RuntimeError("Error")
        out_p = np.empty((m, 3), dtype=np.float32)
        out_f = np.empty((m, fdim), dtype=np.float32) if use_feature else None
        out_c = np.empty((m, ldim), dtype=np.int32) if use_classes else None
        keys = np.empty(m, dtype=np.uint64) if return_keys else None
        counts = np.empty(m, dtype=np.int32) if return_keys else None
        _lib.check(L.ssdr_grid_fetch_ex(h, _lib.ptr(out_p), _lib.ptr(out_f), _lib.ptr(out_c), _lib.ptr(keys),
                                        _lib.ptr(counts)))
    finally:
        L.ssdr_grid_free(h)
    if use_feature and use_classes:
        ret = (out_p, out_f, out_c)
    elif use_feature:
        ret = (out_p, out_f)
    elif use_classes:
        ret = (out_p, out_c)
    else:
        ret = out_p
    if return_keys:
        return ret, keys, counts
    return ret
