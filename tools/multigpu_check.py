"""Run under torchrun on >= 2 GPUs: row-sharded FPS (NCCL 8-byte max all-reduce per pick) must return exactly the
single-GPU picks on every rank; KNN batch items sharded over ranks must equal the single-GPU result; one cloud
subsampled in voxel-layer slabs over the ranks must concatenate to the single-GPU rows.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from ssdr_al_b200 import device as D
from ssdr_al_b200 import dist as SD


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = SD.NcclComm.from_torch(dev)
    ok = True
    for (n, d, picks) in ((200_000, 32, 300), (50_000, 256, 100), (10_007, 32, 64)):
        g = torch.Generator(device=dev)
        g.manual_seed(1234)  # same matrix on every rank
        F = torch.randn((n, d), generator=g, device=dev, dtype=torch.float32)
        want = D.fps(F, picks, 17)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        got = SD.fps_sharded(F, picks, 17, comm)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = bool(torch.equal(got, want))
        ok &= same
        if rank == 0:
            print("fps_sharded N=%d D=%d picks=%d world=%d equal_to_single_gpu=%s  %.1f us/pick" %
                  (n, d, picks, world, same, 1e6 * dt / picks), flush=True)
    # KNN: batch items round-robin over ranks, gathered for comparison
    rng = np.random.default_rng(3)
    pts = torch.from_numpy(rng.random((8, 20000, 3), dtype=np.float32)).to(dev)
    full = D.knn_batch(pts, pts, 16)
    mine = SD.shard_items(8, world, rank)
    part = D.knn_batch(pts[mine].contiguous(), pts[mine].contiguous(), 16)
    same = bool(torch.equal(part, full[mine]))
    ok &= same
    # one cloud over all ranks: voxel-layer slabs, (a) replicated input, (b) row chunks + all-to-all routing
    n = 2_000_000
    p = rng.random((n, 3), dtype=np.float32) * np.array([60.0, 40.0, 8.0], np.float32)
    p[: n // 3, 2] = 1.0 + 0.01 * rng.standard_normal(n // 3).astype(np.float32)
    tp = torch.from_numpy(p).to(dev)
    tf = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).to(dev)
    tc = torch.from_numpy(rng.integers(0, 13, (n, 1)).astype(np.int32)).to(dev)
    want = D.grid_subsample(tp, tf, tc, 0.08, return_keys=True)
    grid_same = []
    for replicated in (True, False):
        b, e = (0, n) if replicated else SD.shard_range(n, world, rank)
        for _ in range(2):  # the second, warm call is the one timed (the first pays NCCL channel set-up)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            got = SD.grid_subsample_sharded(tp[b:e], tf[b:e], tc[b:e], 0.08, replicated=replicated, return_keys=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = got[0].shape[0]
        dist.all_reduce(sizes)
        off = int(sizes[:rank].sum())
        m = got[0].shape[0]
        gsame = int(sizes.sum()) == want[0].shape[0] and all(
            torch.equal(got[j], want[j][off:off + m]) for j in range(3)) and np.array_equal(
                got[3], want[3][off:off + m]) and np.array_equal(got[4], want[4][off:off + m])
        ok &= bool(gsame)
        grid_same.append((replicated, gsame, sizes.tolist(), dt))
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("knn_batch sharded by item equal_to_single_gpu=%s" % same, flush=True)
        for replicated, gs, sizes, dt in grid_same:
            print("grid_subsample_sharded N=%d replicated=%s rows_per_rank=%s equal_to_single_gpu(rank0)=%s  %.1f ms"
                  % (n, replicated, sizes, gs, 1e3 * dt), flush=True)
        print("MULTIGPU_CHECK", "OK" if int(flags.item()) == 1 else "FAILED", flush=True)
    comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flags.item()) == 1 else 1)


if __name__ == "__main__":
    main()
