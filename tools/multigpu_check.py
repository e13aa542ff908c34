"""Equality of every sharded multi-GPU path with its single-GPU result, bit for bit.  Run under torchrun on >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_check.py
or in-process from bench.py (`run_checks`), which fails the bench run on a mismatch.

  * row-sharded FPS / k-center (float32 and float64) with the per-pick exchange fused into the persistent kernel over
    peer memory (PeerGroup; falls back to the per-pick NCCL all-reduce if peer memory cannot be mapped)
  * knn_batch with the batch items sharded over ranks, and ONE cloud with its queries sharded by block (knn_sharded)
  * one cloud subsampled in voxel-layer slabs over the ranks: (a) replicated input, (b) row chunks + all-to-all routing
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


def run_checks(dev, rank, world, comm, log=None):
    """Returns (all_equal, list of result dicts).  `comm`: a SD.PeerGroup or SD.NcclComm."""
    results = []
    ok = True
    p2p = isinstance(comm, SD.PeerGroup)

    def note(name, same, **kw):
        nonlocal ok
        ok &= bool(same)
        r = dict(check=name, equal=bool(same), **kw)
        results.append(r)
        if log and rank == 0:
            log(r)

    cases = [(200_000, 32, torch.float32, 300), (50_000, 256, torch.float32, 100), (10_007, 32, torch.float32, 64)]
    if p2p:
        cases += [(30_000, 129, torch.float64, 80), (40_000, 64, torch.float32, 60)]
    for (n, d, dt, picks) in cases:
        g = torch.Generator(device=dev)
        g.manual_seed(1234)  # same matrix on every rank
        F = torch.randn((n, d), generator=g, device=dev, dtype=dt)
        want = D.fps(F, picks, 17)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        got = SD.fps_sharded(F, picks, 17, comm)
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        note("fps_sharded", torch.equal(got, want), N=n, D=d, dtype=str(dt).split(".")[-1], picks=picks, world=world,
             transport="peer-memory mailboxes" if p2p else "nccl all-reduce per pick", us_per_pick=1e6 * dt_s / picks)
        if p2p:
            sel = torch.arange(n - 25, n, device=dev, dtype=torch.int64)
            want = D.kcenter(F, sel, picks)
            got = SD.kcenter_sharded(F, sel, picks, comm)
            note("kcenter_sharded", torch.equal(got, want), N=n, D=d, dtype=str(dt).split(".")[-1], picks=picks)
    # KNN: batch items round-robin over ranks
    rng = np.random.default_rng(3)
    pts = torch.from_numpy(rng.random((8, 20000, 3), dtype=np.float32)).to(dev)
    full = D.knn_batch(pts, pts, 16)
    mine = SD.shard_items(8, world, rank)
    if mine:
        part = D.knn_batch(pts[mine].contiguous(), pts[mine].contiguous(), 16)
        note("knn_batch items", torch.equal(part, full[mine]), items=len(mine))
    # KNN: one cloud (with duplicated points, so the tie path runs), queries sharded by block, gathered
    cloud = rng.random((300_000, 3), dtype=np.float32) * np.array([30.0, 20.0, 3.0], np.float32)
    cloud[1000:1400] = cloud[:400]
    tc = torch.from_numpy(cloud).to(dev)
    want = D.knn_batch(tc[None], tc[None], 16)[0]
    got, (qb, qe) = SD.knn_sharded(tc, tc, 16, gather=True)
    note("knn_sharded", torch.equal(got, want), N=cloud.shape[0], block=[qb, qe])
    # one cloud over all ranks: voxel-layer slabs, (a) replicated input, (b) row chunks + all-to-all routing
    n = 2_000_000
    p = rng.random((n, 3), dtype=np.float32) * np.array([60.0, 40.0, 8.0], np.float32)
    p[: n // 3, 2] = 1.0 + 0.01 * rng.standard_normal(n // 3).astype(np.float32)
    tp = torch.from_numpy(p).to(dev)
    tf = torch.from_numpy(rng.random((n, 3), dtype=np.float32)).to(dev)
    tcl = torch.from_numpy(rng.integers(0, 13, (n, 1)).astype(np.int32)).to(dev)
    want = D.grid_subsample(tp, tf, tcl, 0.08, return_keys=True)
    for replicated in (True, False):
        b, e = (0, n) if replicated else SD.shard_range(n, world, rank)
        for _ in range(2):  # the second, warm call is the one timed (the first pays NCCL channel set-up)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            got = SD.grid_subsample_sharded(tp[b:e], tf[b:e], tcl[b:e], 0.08, replicated=replicated, return_keys=True)
            torch.cuda.synchronize()
            dt_s = time.perf_counter() - t0
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = got[0].shape[0]
        dist.all_reduce(sizes)
        off = int(sizes[:rank].sum())
        m = got[0].shape[0]
        same = int(sizes.sum()) == want[0].shape[0] and all(
            torch.equal(got[j], want[j][off:off + m]) for j in range(3)) and np.array_equal(
                got[3], want[3][off:off + m]) and np.array_equal(got[4], want[4][off:off + m])
        note("grid_subsample_sharded", same, N=n, replicated=replicated, rows_per_rank=sizes.tolist(), ms=1e3 * dt_s)
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    return int(flags.item()) == 1, results


def make_comm(dev):
    """Peer-memory group when CUDA IPC works between the ranks, else the NCCL communicator."""
    comm = SD.PeerGroup.from_torch(dev)
    return comm if comm is not None else SD.NcclComm.from_torch(dev)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = make_comm(dev)
    ok, _ = run_checks(dev, rank, world, comm, log=lambda r: print(r, flush=True))
    if rank == 0:
        print("MULTIGPU_CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
