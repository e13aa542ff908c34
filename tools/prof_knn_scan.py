"""k=16 KNN of the barycentres of the config-3 scan (12.2 M points, one cloud) -- the target of the ncu capture of the
query kernel on a cloud larger than the L2.    python tools/prof_knn_scan.py [scan points] [calls]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ssdr_al_b200 import device as D
from tools import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x, f, c = synth.scan_cloud(n, 2, torch.device("cuda", 0))
sp = D.grid_subsample(x, None, None, 0.06)[0].contiguous()
del x, f, c
for _ in range(calls):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _, st = D.knn_batch(sp[None], sp[None], 16, want_stats=True)
    b.record()
    torch.cuda.synchronize()
    print("M=%d %.3f ms main %.3f evals/q %.0f" % (sp.shape[0], a.elapsed_time(b), st["main_kernel_ms"],
                                                   st["dist_evals"] / sp.shape[0]), flush=True)
