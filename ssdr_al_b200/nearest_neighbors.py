"""KNN behind the reference's `nearest_neighbors` module interface (utils/nearest_neighbors/knn.pyx:33-149)."""
import numpy as np

from . import _lib


def _check_dim(dim):
    if dim != 3:
        raise RuntimeError("ssdr_al_b200.nearest_neighbors supports dim == 3 only (got %d); every caller in the "
                           "reference passes xyz" % dim)


def knn(pts, queries, K, omp=False):
    """knn.pyx:33-69.  `omp` only selected threading in the reference; results are identical, so it is ignored."""
    pts_c = np.ascontiguousarray(pts, dtype=np.float32)
    queries_c = np.ascontiguousarray(queries, dtype=np.float32)
    _check_dim(pts_c.shape[1])
    # knn.pyx:53 allocates np.zeros; every slot is overwritten unless K > npts, so pinned memory is only zeroed then
    indices = (_lib.pinned_zeros if K > pts_c.shape[0] else _lib.pinned_empty)((queries_c.shape[0], K), np.int64)
    _lib.check(_lib.lib().ssdr_knn(_lib.ptr(pts_c), pts_c.shape[0], pts_c.shape[1], _lib.ptr(queries_c),
                                   queries_c.shape[0], int(K), _lib.ptr(indices)))
    return indices


def _batch_view(a):
    """(array, item stride in floats) for a (B, n, 3) float32 array whose items are dense but possibly spaced out --
    a slice `x[:, :n, :]` of a C-contiguous array -- so that it can be uploaded without packing it on the host first;
    anything else goes through np.ascontiguousarray like in knn.pyx:95-96."""
    if (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 3 and a.shape[0] > 0 and a.shape[1] > 0
            and a.strides[2] == 4 and a.strides[1] == 4 * a.shape[2] and a.strides[0] % 4 == 0
            and a.strides[0] >= a.shape[1] * a.shape[2] * 4):
        return a, a.strides[0] // 4
    c = np.ascontiguousarray(a, dtype=np.float32)
    return c, (c.shape[1] * c.shape[2] if c.ndim == 3 else 0)


def knn_batch(pts, queries, K, omp=False):
    """knn.pyx:71-109."""
    pts_c, ps = _batch_view(pts)
    queries_c, qs = (pts_c, ps) if queries is pts else _batch_view(queries)
    _check_dim(pts_c.shape[2])
    indices = (_lib.pinned_zeros if K > pts_c.shape[1] else _lib.pinned_empty)(
        (pts_c.shape[0], queries_c.shape[1], K), np.int64)
    _lib.check(_lib.lib().ssdr_knn_batch_strided(_lib.ptr(pts_c), pts_c.shape[0], pts_c.shape[1], pts_c.shape[2], ps,
                                                 _lib.ptr(queries_c), queries_c.shape[1], qs, int(K),
                                                 _lib.ptr(indices)))
    return indices


def knn_pyramid(batch_xyz, ratios, K):
    """RandLA-Net's input pyramid, the loop of s3dis_dataset.py:164-177 (tf_map) / helper_tool.py:173-183, in ONE call:

        for i in range(num_layers):
            neigh[i] = knn_batch(xyz, xyz, K);  sub = xyz[:, :N // ratios[i], :];  up[i] = knn_batch(sub, xyz, 1);  xyz = sub

    batch_xyz (B, N, 3) -> (neigh, up): lists of int64 arrays, neigh[i] (B, N_i, K) and up[i] (B, N_i, 1), identical to
    what the ten calls return.  The points travel to the device once, the support clouds are searched side by side
    and every level's rows are copied back while the others still compute."""
    import ctypes as C
    xyz = np.ascontiguousarray(batch_xyz, dtype=np.float32)
    _check_dim(xyz.shape[2])
    B, n = xyz.shape[0], xyz.shape[1]
    L = len(ratios)
    neigh, up = [], []
    for r in ratios:
        neigh.append(_lib.pinned_empty((B, n, K), np.int64))
        up.append(_lib.pinned_empty((B, n, 1), np.int64))
        n //= int(r)
    rat = (C.c_int32 * L)(*[int(r) for r in ratios])
    pn = (C.c_void_p * L)(*[_lib.ptr(a) for a in neigh])
    pu = (C.c_void_p * L)(*[_lib.ptr(a) for a in up])
    _lib.check(_lib.lib().ssdr_knn_pyramid(_lib.ptr(xyz), B, xyz.shape[1], xyz.shape[2], rat, L, int(K), pn, pu))
    return neigh, up


def knn_batch_distance_pick(pts, nqueries, K, omp=False):
    """knn.pyx:111-149.  The reference seeds this with time(0) (knn_.cxx:143), so it is not reproducible, and no
    caller exists anywhere in SSDR-AL; it is exported only so that attribute lookups do not fail."""
    raise NotImplementedError("knn_batch_distance_pick is unused by SSDR-AL and non-deterministic in the reference")
