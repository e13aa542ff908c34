"""Step timing of dist.grid_subsample_sharded on the config-3 scan (torchrun, N ranks): wall clock per step, rank 0.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/prof_sharded_scan.py [points]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ssdr_al_b200 import device as D
from ssdr_al_b200 import dist as SD
from tools import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
xyz, rgb, lab = synth.scan_cloud(n, 2, dev)
lab2 = lab[:, None].contiguous()


def tick(tag, t0, log):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    log.append("%s %.3f" % (tag, 1e3 * (t1 - t0)))
    return t1


for rep in range(4):
    torch.cuda.synchronize()
    dist.barrier()
    log = []
    t0 = time.perf_counter()
    b, e = SD.shard_range(n, world, rank)
    box = torch.tensor(D.grid_bbox(xyz[b:e]), dtype=torch.float32, device=dev)
    t = tick("bbox", t0, log)
    box[3:] = -box[3:]
    dist.all_reduce(box, op=dist.ReduceOp.MIN)
    box[3:] = -box[3:]
    bbox = [float(v) for v in box.cpu()]
    t = tick("allreduce+cpu", t, log)
    axis, bounds = SD.choose_slabs(xyz[b:e], 0.06, bbox, world, "auto", False)
    t = tick("choose_slabs", t, log)
    slab = (axis, int(bounds[rank]), int(bounds[rank + 1]))
    sp, sf, sc = D.grid_subsample(xyz, rgb, lab2, 0.06, bbox=bbox, slab=slab)
    t = tick("slab subsample", t, log)
    full, span = SD.gather_rows(sp)
    t = tick("gather", t, log)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    got = SD.grid_subsample_sharded(xyz, rgb, lab2, 0.06, replicated=True, axis="auto")
    t = tick("whole call", t0, log)
    if rank == 0:
        print("rep %d: %s | rows %d of %d" % (rep, "  ".join(log), sp.shape[0], full.shape[0]), flush=True)
dist.barrier()
dist.destroy_process_group()
