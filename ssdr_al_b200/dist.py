"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* KNN / grid subsampling: batch items, rooms and scans are independent units -> `shard_items` / `shard_range` give
  each rank its units; there is no data-path collective.
* FPS over row shards: every rank scans rows [begin, end) of its full copy of the feature matrix; after each pick
  the packed (distance bits << 32 | ~row) candidates are max-all-reduced (8 bytes) -- `ssdr_fps_f32_sharded` does
  it with NCCL on the device; `pack_candidate` / `unpack_candidate` define the key so that MAX == "largest distance,
  lowest row", i.e. np.argmax semantics across ranks.
* One large scan over several GPUs (`grid_subsample_sharded`): voxel-layer slabs along z with the grid geometry of
  the whole cloud; every voxel is owned by one rank and sees its points in input order, so the rows are bit-identical
  to a single-GPU run and the ranks' outputs concatenated in rank order ARE the single-GPU key-ordered result.  Row
  chunks are routed to their slab owner with one all-to-all (the only exchange step of the path); a cloud that is
  already replicated on every rank needs no exchange at all.
"""
import ctypes as C

import numpy as np

from . import _lib


def shard_range(n, world, rank):
    """Contiguous, balanced [begin, end) of n rows for `rank` (first n % world ranks get one extra row)."""
    base, extra = divmod(int(n), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_items(n_items, world, rank):
    """Round-robin assignment of independent units (batch items, rooms, scans)."""
    return list(range(rank, int(n_items), int(world)))


def pack_candidate(dist, row):
    """uint64 key: float32 distance bits (non-negative, so order preserving) in the high word, ~row in the low."""
    d = np.asarray(dist, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = np.uint64(0xFFFFFFFF) - np.asarray(row, dtype=np.uint64)
    return (d << np.uint64(32)) | r


def unpack_candidate(key):
    key = np.asarray(key, dtype=np.uint64)
    row = np.uint64(0xFFFFFFFF) - (key & np.uint64(0xFFFFFFFF))
    dist = (key >> np.uint64(32)).astype(np.uint32).view(np.float32)
    return dist, row.astype(np.int64)


class NcclComm(object):
    """ncclComm_t created through the library's dlopen'ed NCCL; the unique id travels over torch.distributed."""

    def __init__(self, handle):
        self.handle = handle

    @classmethod
    def from_torch(cls, device):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            _lib.check(_lib.lib().ssdr_nccl_unique_id(_lib.ptr(ident)))
        t = torch.from_numpy(ident).to(device)
        dist.broadcast(t, src=0)
        ident = t.cpu().numpy()
        h = C.c_void_p()
        _lib.check(_lib.lib().ssdr_nccl_comm_init(C.byref(h), world, _lib.ptr(ident), rank))
        return cls(h)

    def destroy(self):
        if self.handle:
            _lib.lib().ssdr_nccl_comm_destroy(self.handle)
            self.handle = None


class PeerGroup(object):
    """The ranks of one node whose persistent selection kernels exchange each pick through peer memory (NVLink):
    every rank owns a small mailbox block, all blocks are mapped into every process with CUDA IPC (include/ssdr_b200.h,
    "peer groups").  `local(world)` builds `world` virtual ranks inside ONE process (tests on a single GPU)."""

    def __init__(self, handle, world, rank):
        self.handle, self.world, self.rank = handle, world, rank

    @classmethod
    def from_torch(cls, device, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.ssdr_peer_group_create(world, rank, C.byref(h)))
        mine = np.zeros(64, np.uint8)
        _lib.check(L.ssdr_peer_group_export(h, _lib.ptr(mine)))
        cuda = dist.get_backend(group) == "nccl"
        t = torch.from_numpy(mine)
        t = t.to(device) if cuda else t
        every = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(every, t, group=group)
        handles = np.ascontiguousarray(np.stack([e.cpu().numpy() for e in every]))
        try:
            _lib.check(L.ssdr_peer_group_connect(h, _lib.ptr(handles)))
            ok = 1
        except RuntimeError:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device if cuda else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # also the barrier that orders creation before use
        if int(flag.item()) == 0:
            L.ssdr_peer_group_destroy(h)
            return None
        return cls(h, world, rank)

    @classmethod
    def local(cls, world):
        L = _lib.lib()
        hs = (C.c_void_p * world)()
        for r in range(world):
            h = C.c_void_p()
            _lib.check(L.ssdr_peer_group_create(world, r, C.byref(h)))
            hs[r] = h
        _lib.check(L.ssdr_peer_group_connect_local(hs, world))
        return [cls(C.c_void_p(hs[r]), world, r) for r in range(world)]

    def check(self):
        """With SSDR_PEER_DEFER_CHECK=1: wait for the current stream and raise if a peer timed out."""
        import torch
        _lib.check(_lib.lib().ssdr_peer_group_check(self.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def destroy(self):
        if self.handle:
            _lib.lib().ssdr_peer_group_destroy(self.handle)
            self.handle = None


def _rows(N, rows, world, rank):
    return rows if rows is not None else shard_range(N, world, rank)


def fps_sharded(F, n_samples, first, comm, rows=None, max_ctas=0):
    """Row-sharded FPS: F is the FULL (N, D) float32/float64 cuda tensor on every rank; returns (n_samples,) int32
    picks, identical on all ranks and equal to the single-GPU picks.  `comm` is a PeerGroup (the per-pick exchange runs
    inside the persistent kernel over NVLink peer memory) or an NcclComm (one launch + 8-byte all-reduce per pick;
    float32 only)."""
    import torch
    F = F.contiguous()
    N = F.shape[0]
    out = torch.zeros(n_samples, dtype=torch.int32, device=F.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if isinstance(comm, PeerGroup):
        begin, end = _rows(N, rows, comm.world, comm.rank)
        dt = 0 if F.dtype == torch.float32 else 1
        _lib.check(_lib.lib().ssdr_fps_sharded_p2p(dt, C.c_void_p(F.data_ptr()), N, F.shape[1], begin, end, int(first),
                                                   int(n_samples), C.c_void_p(out.data_ptr()), comm.handle, stream,
                                                   int(max_ctas)))
        return out
    import torch.distributed as dist
    begin, end = _rows(N, rows, dist.get_world_size(), dist.get_rank())
    _lib.check(_lib.lib().ssdr_fps_f32_sharded(C.c_void_p(F.data_ptr()), N, F.shape[1], begin, end, int(first),
                                               int(n_samples), C.c_void_p(out.data_ptr()), comm.handle, stream))
    return out


def kcenter_sharded(X, selected, n_pick, comm, rows=None, max_ctas=0):
    """Row-sharded k-center greedy (kCenterGreedy.select_batch_): X is the FULL (N, D) cuda tensor on every rank,
    `selected` the already chosen rows (int64 cuda); returns (n_pick,) int64 picks, identical on all ranks."""
    import torch
    X = X.contiguous()
    N = X.shape[0]
    out = torch.zeros(n_pick, dtype=torch.int64, device=X.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    begin, end = _rows(N, rows, comm.world, comm.rank)
    dt = 0 if X.dtype == torch.float32 else 1
    sel = selected.contiguous() if selected is not None and selected.numel() else None
    _lib.check(_lib.lib().ssdr_kcenter_sharded_p2p(dt, C.c_void_p(X.data_ptr()), N, X.shape[1], begin, end,
                                                   C.c_void_p(sel.data_ptr()) if sel is not None else None,
                                                   sel.numel() if sel is not None else 0, int(n_pick),
                                                   C.c_void_p(out.data_ptr()), comm.handle, stream, int(max_ctas)))
    return out


def knn_sharded(pts, queries, K, group=None, gather=False, int32=False):
    """k-NN of ONE cloud with all ranks: the support cloud `pts` (N,3) (and with it the cell grid built on the device)
    is replicated on every rank, the queries (Q,3) are sharded by contiguous block -- the loop the reference
    parallelises with OpenMP (knn_.cxx:57-58).  Returns (rows of this rank's block (q_end-q_begin, K), (q_begin, q_end)),
    or with gather=True the full (Q, K) result on every rank (one all-gather of the index rows)."""
    import torch
    import torch.distributed as dist
    from . import device as dev
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Q = queries.shape[0]
    b, e = shard_range(Q, world, rank)
    if e > b:
        mine = dev.knn_batch(pts[None], queries[None, b:e], K, int32=int32)[0]
    else:
        mine = torch.zeros((0, K), dtype=torch.int32 if int32 else torch.int64, device=pts.device)
    if not gather:
        return mine, (b, e)
    spans = [shard_range(Q, world, r) for r in range(world)]
    width = max(s[1] - s[0] for s in spans)  # equal-sized pieces for the collective (blocks differ by at most one row)
    padded = torch.zeros((width, K), dtype=mine.dtype, device=pts.device)
    padded[: e - b] = mine
    every = torch.empty((world, width, K), dtype=mine.dtype, device=pts.device)
    dist.all_gather_into_tensor(every, padded, group=group)
    return torch.cat([every[r, : s[1] - s[0]] for r, s in enumerate(spans)]), (b, e)


def gather_rows(rows, group=None):
    """All-gather row blocks of different lengths (the slabs of a sharded subsampling) into the full array, rank
    order == row order.  Returns (full (sum m_r, ...), (begin, end) of this rank's block).  Over NCCL every block
    travels once, straight into its place in the result (a group of broadcasts, no padding and no re-packing); other
    backends (gloo in the CPU tests) and empty blocks take the padded all-gather."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=rows.device)
    sizes[rank] = rows.shape[0]
    dist.all_reduce(sizes, group=group)
    sizes = sizes.tolist()
    begin = sum(sizes[:rank])
    span = (begin, begin + sizes[rank])
    if rows.is_cuda and min(sizes) > 0 and dist.get_backend(group) == "nccl":
        full = torch.empty((sum(sizes),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        at, parts = 0, []
        for n in sizes:
            parts.append(full[at:at + n])
            at += n
        dist.all_gather(parts, rows.contiguous(), group=group)
        return full, span
    width = max(max(sizes), 1)
    padded = torch.zeros((width,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    # (output in the concatenated form: gloo accepts no other, NCCL accepts both)
    every = torch.empty((world * width,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(every, padded, group=group)
    full = torch.cat([every[r * width: r * width + sizes[r]] for r in range(world)])
    return full, span


def balanced_slabs(layer_counts, world):
    """Cut n_layers voxel layers into `world` contiguous slabs with near-equal point counts.

    Returns bounds (world + 1,) int64 with bounds[0] = 0, bounds[-1] = n_layers, non-decreasing: rank r owns layers
    [bounds[r], bounds[r+1]).  Deterministic in `layer_counts`, so every rank derives the same cuts."""
    cnt = np.asarray(layer_counts, dtype=np.int64)
    n_layers, world = int(cnt.shape[0]), int(world)
    cum = np.cumsum(cnt)
    total = int(cum[-1]) if n_layers else 0
    bounds = np.zeros(world + 1, np.int64)
    bounds[-1] = n_layers
    for r in range(1, world):
        target = (total * r + world - 1) // world  # first layer boundary with at least r/world of the points below
        bounds[r] = int(np.searchsorted(cum, target, side="left")) + 1 if total else 0
    bounds[1:-1] = np.minimum(bounds[1:-1], n_layers)
    return np.maximum.accumulate(bounds)


def route_plan(layers, bounds):
    """Owner rank of each point from its layer index; returns (dest int64 array, per-rank send counts)."""
    layers = np.asarray(layers, dtype=np.int64)
    bounds = np.asarray(bounds, dtype=np.int64)
    dest = np.searchsorted(bounds[1:-1], layers, side="right")
    return dest, np.bincount(dest, minlength=len(bounds) - 1)


MAX_LAYERS = 1 << 24


def exchange_groups(tensors, send_counts, group=None):
    """One all-to-all of row-aligned tensors whose rows are already grouped by destination rank (send_counts[r] rows
    for rank r, in that order).  The received pieces arrive concatenated in source-rank order, so rows that were in
    global input order inside every group stay in global input order on their owner.  None entries pass through."""
    import torch
    import torch.distributed as dist
    first = next(t for t in tensors if t is not None)
    send = torch.tensor([int(v) for v in send_counts], dtype=torch.int64, device=first.device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = send.tolist(), recv.tolist()

    def exchange(t):
        if t is None:
            return None
        out = torch.empty((sum(recv_l),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(out, t.contiguous(), output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
        return out

    return tuple(exchange(t) for t in tensors)


def slab_exchange(points, features, classes, sampleDl, axis, bbox, bounds, group=None):
    """The exchange step of sharded subsampling: the library groups this rank's rows by slab owner on the device (stable
    partition, input order kept: ssdr_grid_route_dev), one all-to-all moves every group to its owner."""
    from . import device as dev
    p, f, c, send_l = dev.grid_route(points, features, classes, sampleDl, axis, bbox, bounds)
    return exchange_groups((p, f, c), send_l, group=group)


def choose_slabs(points, sampleDl, bbox, world, axis, replicated, group=None, sample_stride=16):
    """(axis, bounds): balanced voxel-layer slabs.  axis = 0/1/2 keeps the caller's axis; "auto" takes the axis whose
    balanced cut has the lightest heaviest slab (a terrestrial scan has almost all of its points in two or three z
    layers, so z slabs cannot be balanced, while x or y slabs can).  Counts come from every sample_stride-th point --
    ownership is by layer range, so the cut positions need not be exact, only identical on every rank.  `points` is
    this rank's row chunk (replicated=False: the histograms of all candidate axes are summed over the ranks in ONE
    all-reduce and read back once) or the whole cloud (replicated=True: no collective)."""
    import torch
    import torch.distributed as dist
    from . import device as dev
    axes, hists = [], []
    for ax in ((0, 1, 2) if axis == "auto" else (int(axis),)):
        n_layers = dev.grid_layers(bbox, sampleDl, ax)
        if n_layers > MAX_LAYERS:
            if axis == "auto":
                continue
            raise ValueError("grid has %d layers along axis %d (limit %d): sampleDl too small for this extent"
                             % (n_layers, ax, MAX_LAYERS))
        axes.append(ax)
        hists.append(dev.grid_layer_hist(points, sampleDl, ax, bbox, sample_stride))
    if not axes:
        raise ValueError("no axis of the grid has at most %d layers: sampleDl too small for this extent" % MAX_LAYERS)
    every = torch.cat(hists) if len(hists) > 1 else hists[0]
    if not replicated:
        dist.all_reduce(every, group=group)
    every = every.cpu().numpy()
    best, at = None, 0
    for ax, hist in zip(axes, hists):
        h = every[at:at + hist.numel()]
        at += hist.numel()
        bounds = balanced_slabs(h, world)
        load = max(int(h[bounds[r]:bounds[r + 1]].sum()) for r in range(world))
        if best is None or load < best[0]:  # ties keep the lower axis: deterministic
            best = (load, ax, bounds)
    return best[1], best[2]


def grid_subsample_sharded(points, features=None, classes=None, sampleDl=0.1, *, replicated=False, group=None,
                           axis=2, return_keys=False, return_axis=False):
    """Subsample ONE cloud with all ranks of `group`; returns this rank's slab of the result as cuda tensors.

    replicated=False: `points` (and features / classes) are this rank's contiguous row chunk of the cloud, chunks in
    rank order.  bbox and layer histogram are all-reduced (a few KB), then every point travels to its slab owner in
    one all-to-all that keeps input order.  replicated=True: every rank holds the whole cloud; no exchange.
    Every voxel has exactly one owner that sees its points in input order, so every row carries the bits of the
    single-GPU run.  axis=2 (default): the ranks' rows concatenated in rank order ARE `device.grid_subsample` of the
    whole cloud (keys are z-major).  axis="auto" picks the best balanced slab axis (see choose_slabs); the union of the
    ranks' rows is then the same set of rows, in rank-major instead of key order (return_keys gives the keys)."""
    import torch
    import torch.distributed as dist
    from . import device as dev
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = points.shape[0]
    if classes is not None and classes.dim() == 1:
        classes = classes[:, None]
    # 1. geometry of the whole cloud.  A replicated cloud is scanned in row chunks too (every rank reads 1/world of it
    # for the corners and for the layer histograms; the few KB of partial results are all-reduced).
    mine = points
    if replicated and world > 1:
        b, e = shard_range(n_local, world, rank)
        mine = points[b:e]
    if mine.shape[0]:
        box = torch.tensor(dev.grid_bbox(mine), dtype=torch.float32, device=points.device)
    else:
        box = torch.tensor([float("inf")] * 3 + [float("-inf")] * 3, dtype=torch.float32, device=points.device)
    if world > 1:
        box[3:] = -box[3:]  # one MIN all-reduce for both corners
        dist.all_reduce(box, op=dist.ReduceOp.MIN, group=group)
        box[3:] = -box[3:]
    bbox = [float(v) for v in box.cpu()]
    # 2. balanced slabs from the (sampled) layer histogram, counted by the library
    axis, bounds = choose_slabs(mine, sampleDl, bbox, world, axis, world == 1, group=group)
    tail = (axis,) if return_axis else ()
    if replicated:
        slab = (axis, int(bounds[rank]), int(bounds[rank + 1]))
        return dev.grid_subsample(points, features, classes, sampleDl, bbox=bbox, slab=slab,
                                  return_keys=return_keys) + tail
    # 3. route every point to its slab owner
    p, f, c = slab_exchange(points, features, classes, sampleDl, axis, bbox, bounds, group=group)
    if p.shape[0] == 0:
        e = lambda t, dt: None if t is None else torch.empty((0, t.shape[1]), dtype=dt, device=points.device)
        res = (torch.empty((0, 3), dtype=torch.float32, device=points.device), e(features, torch.float32),
               e(classes, torch.int32))
        return (res + (np.empty(0, np.uint64), np.empty(0, np.int32)) if return_keys else res) + tail
    return dev.grid_subsample(p, f, c, sampleDl, bbox=bbox, slab=None, return_keys=return_keys) + tail
