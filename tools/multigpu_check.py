"""Run under torchrun on >= 2 GPUs: row-sharded FPS (NCCL 8-byte max all-reduce per pick) must return exactly the
single-GPU picks on every rank; KNN batch items sharded over ranks must equal the single-GPU result.
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
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("knn_batch sharded by item equal_to_single_gpu=%s" % same, flush=True)
        print("MULTIGPU_CHECK", "OK" if int(flags.item()) == 1 else "FAILED", flush=True)
    comm.destroy()
    dist.destroy_process_group()
    sys.exit(0 if int(flags.item()) == 1 else 1)


if __name__ == "__main__":
    main()
