"""Config 3, the KNN half: k=16 of the barycentres of a scan, whole (one GPU) and as the query block rank r of `world`
ranks would run (support cloud, cell grid and tie-path trees over the WHOLE cloud), with the stage breakdown.
    python tools/prof_cfg3_knn.py [scan points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ssdr_al_b200 import device as D
from ssdr_al_b200 import dist as SD
from tools import synth

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
x, f, c = synth.scan_cloud(n, 2, dev)
sp = D.grid_subsample(x, None, None, 0.06)[0].contiguous()
del x, f, c
M = sp.shape[0]
print("scan %d points -> %d barycentres" % (n, M), flush=True)
for world in (1, 2, 8):
    for r in sorted({0, world // 2, world - 1}):
        b, e = SD.shard_range(M, world, r)
        q = sp[None, b:e].contiguous()
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, st = D.knn_batch(sp[None], q, 16, want_stats=True)
            e1.record()
            torch.cuda.synchronize()
            print("world %d rank %d call %d: %.3f ms | grid %.3f main %.3f tie %.3f (tree %.3f) rows %d builds %d" % (
                world, r, rep, e0.elapsed_time(e1), st["grid_build_ms"], st["main_kernel_ms"], st["tie_path_ms"],
                st["tree_build_ms"], st["tie_rows"], st["tree_builds"]), flush=True)
