"""Phase timing of grid subsampling (device %globaltimer marks of the sort / reduce kernels) + CUDA-event totals."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ssdr_al_b200 import _lib, device as D
from tools import synth

dev = torch.device("cuda", 0)
NAMES = ["geometry", "keys", "pass0", "pass1", "pass2", "pass3", "pass4", "pass5", "pass6", "pass7", "heads", "starts",
         "gap->reduce", "reduce"]


def report(tag, pts, f, c, dl):
    for _ in range(3):
        D.grid_subsample(pts, f, c, dl)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = D.grid_subsample(pts, f, c, dl)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    m = (C.c_uint64 * 16)()
    kb = C.c_int(0)
    _lib.lib().ssdr_grid_debug_timing(m, C.byref(kb))
    m = list(m)
    parts = []
    prev = m[0]
    for i, name in enumerate(NAMES):
        v = m[i + 1]
        if v:
            parts.append("%s %.1f" % (name, (v - prev) / 1e3))
            prev = v
    print("%s: N=%d M=%d key_bits=%d  event total %.3f ms (min %.3f)  device span %.1f us | %s" % (
        tag, pts.shape[0], r[0].shape[0], kb.value, float(np.median(ts)), min(ts), (m[14] - m[0]) / 1e3, " ".join(parts)), flush=True)


p, rgb, lab = synth.room_cloud(1_000_000, 0)
report("room 1M", torch.from_numpy(p).to(dev), torch.from_numpy(rgb.astype(np.float32)).to(dev),
       torch.from_numpy(lab.astype(np.int32)).to(dev), 0.04)
report("room 1M xyz only", torch.from_numpy(p).to(dev), None, None, 0.04)
rng = np.random.default_rng(0)
u = torch.from_numpy((rng.random((1_000_000, 3)) * np.array([7.0, 5.0, 3.0])).astype(np.float32)).to(dev)
report("uniform 1M", u, torch.from_numpy(rgb.astype(np.float32)).to(dev), torch.from_numpy(lab.astype(np.int32)).to(dev), 0.04)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
x, f, c = synth.scan_cloud(n, 2, dev)
report("scan %dM" % (n // 1_000_000), x, f, c[:, None].contiguous(), 0.06)

# ---- the multi-GPU slab path on one device: what each of `world` ranks would run on the replicated scan
from ssdr_al_b200 import dist as SD  # noqa: E402


def ev_ms(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), r


c2 = c[:, None].contiguous()
for world in (2, 8):
    t_box, box = ev_ms(lambda: D.grid_bbox(x))
    # the slab axis the sharded path would pick ("auto": lightest heaviest slab by sampled point counts)
    best = None
    t_hist_all = 0.0
    for ax in (0, 1, 2):
        if D.grid_layers(box, 0.06, ax) > SD.MAX_LAYERS:
            continue
        t_hist, hist = ev_ms(lambda: D.grid_layer_hist(x, 0.06, ax, box, 16))
        t_hist_all += t_hist
        h = hist.cpu().numpy()
        bnd = SD.balanced_slabs(h, world)
        load = max(int(h[bnd[r]:bnd[r + 1]].sum()) for r in range(world))
        if best is None or load < best[0]:
            best = (load, ax, bnd, hist.numel())
    _, axis, bounds, nl = best
    print("slab path world=%d: bbox %.3f ms, layer hists %.3f ms, axis %d (%d layers), bounds %s" % (
        world, t_box, t_hist_all, axis, nl, bounds.tolist()), flush=True)
    for r in range(world):
        slab = (axis, int(bounds[r]), int(bounds[r + 1]))
        D.grid_subsample(x, f, c2, 0.06, bbox=box, slab=slab)
        t, res = ev_ms(lambda: D.grid_subsample(x, f, c2, 0.06, bbox=box, slab=slab))
        m = (C.c_uint64 * 16)()
        kb = C.c_int(0)
        _lib.lib().ssdr_grid_debug_timing(m, C.byref(kb))
        m = list(m)
        parts, prev = [], m[0]
        for i, name in enumerate(NAMES):
            v = m[i + 1]
            if v:
                parts.append("%s %.1f" % (name, (v - prev) / 1e3))
                prev = v
        print("  rank %d layers [%d,%d): voxels %d  event %.3f ms | %s" % (r, slab[1], slab[2], res[0].shape[0], t,
                                                                         " ".join(parts)), flush=True)
