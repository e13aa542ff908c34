"""Device-resident entry points: the same C ABI, fed with CUDA tensors (torch is used only for device memory and
streams).  Work is enqueued on torch's current stream."""
import ctypes as C
import threading

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def knn_batch(pts, queries, K, out=None, want_stats=False, int32=False):
    """pts (B,N,3) f32 cuda, queries (B,Q,3) f32 cuda -> (B,Q,K) int64 (or int32) cuda tensor [, stats dict]."""
    assert pts.is_cuda and queries.is_cuda and pts.dtype == torch.float32 and queries.dtype == torch.float32
    pts = pts.contiguous()
    queries = pts if queries is pts else queries.contiguous()
    B, N, _ = pts.shape
    Q = queries.shape[1]
    if out is None:
        out = torch.zeros((B, Q, K), dtype=torch.int32 if int32 else torch.int64, device=pts.device)
    st = _lib.KnnStats() if want_stats else None
    fn = _lib.lib().ssdr_knn_batch_dev_i32 if out.dtype == torch.int32 else _lib.lib().ssdr_knn_batch_dev
    _lib.check(fn(_p(pts), B, N, _p(queries), Q, int(K), _p(out), _stream(), C.byref(st) if st else None))
    if want_stats:
        return out, {f: getattr(st, f) for f, _ in _lib.KnnStats._fields_}
    return out


def knn_pyramid(xyz, ratios, K, neigh=None, up=None, check=False):
    """RandLA-Net's input pyramid (s3dis_dataset.py:164-177) in one asynchronous call: xyz (B,N,3) f32 cuda ->
    (neigh, up) lists with neigh[l] (B,N_l,K) and up[l] (B,N_l,1) int64, N_{l+1} = N_l // ratios[l].  Nothing is read
    back; check=True waits for the stream and raises if the exact tie path reported an error."""
    assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.dim() == 3 and xyz.shape[2] == 3
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    L = len(ratios)
    sizes = [N]
    for r in ratios:
        sizes.append(sizes[-1] // int(r))
    if neigh is None:
        neigh = [torch.zeros((B, sizes[l], K), dtype=torch.int64, device=xyz.device) for l in range(L)]
    if up is None:
        up = [torch.zeros((B, sizes[l], 1), dtype=torch.int64, device=xyz.device) for l in range(L)]
    rat = (C.c_int32 * L)(*[int(r) for r in ratios])
    pn = (C.c_void_p * L)(*[t.data_ptr() for t in neigh])
    pu = (C.c_void_p * L)(*[t.data_ptr() for t in up])
    _lib.check(_lib.lib().ssdr_knn_pyramid_dev(_p(xyz), B, N, rat, L, int(K), pn, pu, _stream()))
    if check:
        _lib.check(_lib.lib().ssdr_knn_status(_stream()))
    return neigh, up


def grid_subsample(points, features=None, classes=None, sampleDl=0.1, bbox=None, slab=None, return_keys=False,
                   compact=None):
    """points (N,3) f32, features (N,fdim) f32, classes (N,ldim) i32 cuda tensors -> tuple of cuda tensors.

    bbox: 6 floats (min xyz, max xyz) of the larger cloud these points belong to (default: their own min/max).
    slab: (axis, layer_lo, layer_hi) -- reduce only the voxel layers [lo, hi) along `axis` (multi-GPU ownership).
    return_keys: also return (keys uint64, counts int32) as numpy arrays.
    The kernels write straight into torch tensors sized for the worst case (one voxel per point: the voxel count is
    only known on the device) and the first M rows are returned -- as views when that wastes little, as compact copies
    when M is much smaller than N (compact=None decides by size; True / False force it).  One stream synchronisation."""
    points = _grid_arg(points, torch.float32, "points", 3)
    features = _grid_arg(features, torch.float32, "features")
    classes = _grid_arg(classes, torch.int32, "classes")
    N = points.shape[0]
    if (features is not None and features.shape[0] != N) or (classes is not None and classes.shape[0] != N):
        raise ValueError("features / classes must have one row per point")
    fdim = features.shape[1] if features is not None else 0
    ldim = (1 if classes.dim() == 1 else classes.shape[1]) if classes is not None else 0
    dev = points.device
    # Worst-case-sized outputs.  A large cloud's come from ONE arena per device that is kept between calls (the rows
    # leave it as compact copies): a fresh 80 M-row allocation per call would make torch's caching allocator split
    # and re-grow its big blocks around the caller's other tensors -- milliseconds of cudaMalloc per call.
    row_bytes = 12 + 4 * fdim + 4 * ldim + (12 if return_keys else 0)
    arena = compact is not False and N * row_bytes > (256 << 20)
    if arena:
        compact = True
        buf = _arena(dev, N * row_bytes + 5 * 256)
        at = [0]

        def carve(shape, dtype):
            n_bytes = _ITEMSIZE[dtype] * int(shape[0]) * (int(shape[1]) if len(shape) > 1 else 1)
            t = buf[at[0]:at[0] + n_bytes].view(dtype).view(shape)
            at[0] += (n_bytes + 255) // 256 * 256
            return t
    else:
        carve = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
    out_k = carve((N,), torch.int64) if return_keys else None  # (8-byte rows first: every piece stays aligned)
    out_p = carve((N, 3), torch.float32)
    out_f = carve((N, fdim), torch.float32) if fdim else None
    out_c = carve((N, ldim), torch.int32) if ldim else None
    out_n = carve((N,), torch.int32) if return_keys else None
    box = (C.c_float * 6)(*[float(v) for v in bbox]) if bbox is not None else None
    axis, lo, hi = slab if slab is not None else (-1, 0, 0)
    M = C.c_size_t(0)
    _lib.check(_lib.lib().ssdr_grid_subsample_into_dev(
        _p(points), _p(features), _p(classes), N, fdim, ldim, float(sampleDl), box, int(axis), int(lo), int(hi),
        _p(out_p), _p(out_f), _p(out_c), _p(out_k), _p(out_n), N, _stream(), C.byref(M)))
    m = M.value
    if compact is None:
        compact = N * (12 + 4 * fdim + 4 * ldim) > (32 << 20) and 2 * m < N
    cut = (lambda t: None if t is None else (t[:m].clone() if compact else t[:m]))
    res = (cut(out_p), cut(out_f), cut(out_c))
    if return_keys:
        import numpy as np
        return res + (out_k[:m].cpu().numpy().view(np.uint64), out_n[:m].cpu().numpy())
    return res


_ITEMSIZE = {torch.float32: 4, torch.int32: 4, torch.int64: 8}
_ARENA = threading.local()  # per calling thread, like the library's own workspaces


def _arena(dev, n_bytes):
    """Grow-only scratch of the calling thread on `dev` (uint8 tensor of at least n_bytes).  Calls that share it must
    be ordered on one stream (they are when the thread keeps torch's current stream)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    pool = _ARENA.__dict__.setdefault("pool", {})
    buf = pool.get(key)
    if buf is None or buf.numel() < n_bytes:
        pool[key] = None  # drop the old block before the larger one is allocated
        buf = pool[key] = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    return buf


def release_scratch():
    """Give the calling thread's output arenas of large grid_subsample calls back to torch's allocator."""
    _ARENA.__dict__.pop("pool", None)


def _grid_arg(t, dtype, name, width=None):
    """The C ABI reads raw device memory: refuse anything that is not a contiguous CUDA tensor of the stated dtype
    instead of reinterpreting its bytes (torch's default int64 labels, float64 points, strided views ...)."""
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s (convert explicitly: the library does not cast)" % (name, dtype, t.dtype))
    if width is not None and (t.dim() != 2 or t.shape[1] != width):
        raise ValueError("%s must have shape (N, %d)" % (name, width))
    if width is None and t.dim() not in (1, 2):
        raise ValueError("%s must be 1-D or 2-D" % name)
    return t.contiguous()


def grid_bbox(points):
    """(N,3) f32 cuda -> [minx, miny, minz, maxx, maxy, maxz] (python floats holding float32 values)."""
    points = _grid_arg(points, torch.float32, "points", 3)
    box = (C.c_float * 6)()
    _lib.check(_lib.lib().ssdr_grid_bbox_dev(_p(points), points.shape[0], _stream(), box))
    return [float(v) for v in box]


def grid_point_layers(points, sampleDl, axis, bbox=None):
    """Voxel layer index of every point along `axis` (int32 cuda tensor) and the number of layers of the grid."""
    points = _grid_arg(points, torch.float32, "points", 3)
    out = torch.empty(points.shape[0], dtype=torch.int32, device=points.device)
    box = (C.c_float * 6)(*[float(v) for v in bbox]) if bbox is not None else None
    n_layers = C.c_ulonglong(0)
    _lib.check(_lib.lib().ssdr_grid_point_layers_dev(_p(points), points.shape[0], box, float(sampleDl), int(axis),
                                                     _p(out), _stream(), C.byref(n_layers)))
    return out, int(n_layers.value)


def grid_layers(bbox, sampleDl, axis):
    """Number of voxel layers of the grid along `axis` (the float32 arithmetic of grid_subsampling.cpp:27-31)."""
    import numpy as np
    dl = np.float32(sampleDl)
    mn, mx = np.float32(bbox[axis]), np.float32(bbox[3 + axis])
    origin = np.floor(mn * (np.float32(1) / dl)) * dl
    return int(np.int64(np.floor((mx - origin) / dl))) + 1


def grid_layer_hist(points, sampleDl, axis, bbox, sample_stride=1):
    """Points per voxel layer along `axis` (int64 cuda tensor of grid_layers(...) entries), counted by the library;
    sample_stride = k counts every k-th point only (slab balancing does not need exact counts)."""
    points = _grid_arg(points, torch.float32, "points", 3)
    n_layers = grid_layers(bbox, sampleDl, axis)
    hist = torch.zeros(n_layers, dtype=torch.int64, device=points.device)
    if points.shape[0]:
        box = (C.c_float * 6)(*[float(v) for v in bbox])
        _lib.check(_lib.lib().ssdr_grid_layer_hist_dev(_p(points), points.shape[0], box, float(sampleDl), int(axis),
                                                       _p(hist), n_layers, int(sample_stride), _stream()))
    return hist


def grid_route(points, features, classes, sampleDl, axis, bbox, bounds):
    """Stable device-side partition of the rows by slab owner (rank r owns layers [bounds[r], bounds[r+1])): returns
    (points, features, classes grouped by destination in input order, per-destination row counts as a python list)."""
    points = _grid_arg(points, torch.float32, "points", 3)
    features = _grid_arg(features, torch.float32, "features")
    classes = _grid_arg(classes, torch.int32, "classes")
    world = len(bounds) - 1
    N = points.shape[0]
    fdim = features.shape[1] if features is not None else 0
    ldim = (1 if classes.dim() == 1 else classes.shape[1]) if classes is not None else 0
    op = torch.empty_like(points)
    of = torch.empty_like(features) if features is not None else None
    oc = torch.empty_like(classes) if classes is not None else None
    counts = (C.c_ulonglong * world)()
    if N:
        box = (C.c_float * 6)(*[float(v) for v in bbox])
        bnd = (C.c_ulonglong * (world + 1))(*[int(b) for b in bounds])
        _lib.check(_lib.lib().ssdr_grid_route_dev(_p(points), _p(features), _p(classes), N, fdim, ldim, float(sampleDl),
                                                  box, int(axis), bnd, world, _p(op), _p(of), _p(oc), counts, _stream()))
    return op, of, oc, [int(v) for v in counts]


_cudart = None


def _copy_d2d(dst, src_ptr, nbytes):
    global _cudart
    if _cudart is None:
        _cudart = C.CDLL("libcudart.so.12")
    rc = _cudart.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), src_ptr, C.c_size_t(nbytes), C.c_int(3), _stream())
    if rc != 0:
        raise RuntimeError("cudaMemcpyAsync failed: %d" % rc)


def fps(F, n_samples, first, out=None):
    """F (N,D) f32/f64 cuda -> (n_samples,) int32 cuda."""
    F = F.contiguous()
    if out is None:
        out = torch.zeros(n_samples, dtype=torch.int32, device=F.device)
    fn = _lib.lib().ssdr_fps_f32_dev if F.dtype == torch.float32 else _lib.lib().ssdr_fps_f64_dev
    _lib.check(fn(_p(F), F.shape[0], F.shape[1], int(first), int(n_samples), _p(out), _stream()))
    return out


def kcenter(X, selected, n_pick, out=None):
    """X (N,D) f32/f64 cuda, selected int64 cuda -> (n_pick,) int64 cuda."""
    X = X.contiguous()
    if selected.dtype != torch.int64 or not selected.is_cuda:
        raise TypeError("selected must be an int64 CUDA tensor")
    selected = selected.contiguous()
    if selected.numel() and (int(selected.min()) < 0 or int(selected.max()) >= X.shape[0]):
        raise ValueError("selected holds a row index outside [0, %d)" % X.shape[0])  # the host entry checks the same
    if out is None:
        out = torch.zeros(n_pick, dtype=torch.int64, device=X.device)
    fn = _lib.lib().ssdr_kcenter_f32_dev if X.dtype == torch.float32 else _lib.lib().ssdr_kcenter_f64_dev
    _lib.check(fn(_p(X), X.shape[0], X.shape[1], _p(selected), selected.numel(), int(n_pick), _p(out), _stream()))
    return out
