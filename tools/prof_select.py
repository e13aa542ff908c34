"""Small driver for profiling the selection (FPS / k-center) persistent kernel and grid subsampling under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ssdr_al_b200 import device as D

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "fps32"
picks = int(sys.argv[2]) if len(sys.argv) > 2 else 200
nrows = int(sys.argv[3]) if len(sys.argv) > 3 else 500_000
g = torch.Generator(device=dev); g.manual_seed(3)
if which.startswith("fps") or which.startswith("kc"):
    d = int(which[3:]) if which.startswith("fps") else int(which[2:])
    F = torch.randn((nrows, d), generator=g, device=dev, dtype=torch.float32)
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if which.startswith("fps"):
            out = D.fps(F, picks, 12345)
        else:
            out = D.kcenter(F, torch.arange(nrows - 16, nrows, device=dev), picks)
        b.record(); torch.cuda.synchronize()
        print(which, "rows", nrows, "picks", picks, "ms", a.elapsed_time(b), "us/pick", 1e3 * a.elapsed_time(b) / picks, flush=True)
elif which == "grid":
    from tools import synth
    n = 1_000_000
    p, rgb_h, lab_h = synth.room_cloud(n, 0)  # BASELINE config 1 (what bench.py's extra.grid_subsample times)
    pts = torch.from_numpy(p).to(dev)
    rgb = torch.from_numpy(rgb_h.astype(np.float32)).to(dev)
    lab = torch.from_numpy(lab_h.astype(np.int32)).to(dev)
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = D.grid_subsample(pts, rgb, lab, 0.04); b.record(); torch.cuda.synchronize()
        print("grid", n, "->", r[0].shape[0], "ms", a.elapsed_time(b), flush=True)
