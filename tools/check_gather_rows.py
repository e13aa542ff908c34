"""dist.gather_rows over NCCL with empty and uneven slabs (both code paths: grouped broadcasts, padded all-gather).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_gather_rows.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
from ssdr_al_b200 import dist as SD
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for sizes in ((0, 5), (4, 0), (3, 7)):
    n = sizes[rank]
    rows = torch.stack([torch.arange(n, dtype=torch.float32, device=dev) + 1000 * rank, torch.full((n,), float(rank), device=dev)], 1)
    full, span = SD.gather_rows(rows)
    want = torch.cat([torch.stack([torch.arange(sizes[r], dtype=torch.float32, device=dev) + 1000 * r, torch.full((sizes[r],), float(r), device=dev)], 1) for r in range(world)])
    ok &= bool(torch.equal(full, want)) and span == (sum(sizes[:rank]), sum(sizes[:rank]) + n)
print("rank", rank, "gather_rows", "OK" if ok else "FAILED", flush=True)
dist.barrier(); dist.destroy_process_group()
