"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* KNN / grid subsampling: batch items, rooms and scans are independent units -> `shard_items` / `shard_range` give
  each rank its units; there is no data-path collective.
* FPS over row shards: every rank scans rows [begin, end) of its full copy of the feature matrix; after each pick
  the packed (distance bits << 32 | ~row) candidates are max-all-reduced (8 bytes) -- `ssdr_fps_f32_sharded` does
  it with NCCL on the device; `pack_candidate` / `unpack_candidate` define the key so that MAX == "largest distance,
  lowest row", i.e. np.argmax semantics across ranks.
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


def fps_sharded(F, n_samples, first, comm, rows=None):
    """Row-sharded FPS: F is the FULL (N, D) float32 cuda tensor on every rank; returns (n_samples,) int32 picks
    (identical on all ranks)."""
    import torch
    import torch.distributed as dist
    F = F.contiguous()
    N = F.shape[0]
    begin, end = rows if rows is not None else shard_range(N, dist.get_world_size(), dist.get_rank())
    out = torch.zeros(n_samples, dtype=torch.int32, device=F.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().ssdr_fps_f32_sharded(C.c_void_p(F.data_ptr()), N, F.shape[1], begin, end, int(first),
                                               int(n_samples), C.c_void_p(out.data_ptr()), comm.handle, stream))
    return out
