"""One Semantic3D-scale scan (config 3) through grid subsampling, a few calls -- the target of the ncu captures of the
large-cloud kernels.    python tools/prof_grid_scan.py [points] [calls]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ssdr_al_b200 import device as D
from tools import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x, f, c = synth.scan_cloud(n, 2, torch.device("cuda", 0))
c2 = c[:, None].contiguous()
for _ in range(calls):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = D.grid_subsample(x, f, c2, 0.06)
    b.record()
    torch.cuda.synchronize()
    print("N=%d M=%d %.3f ms" % (n, r[0].shape[0], a.elapsed_time(b)), flush=True)
