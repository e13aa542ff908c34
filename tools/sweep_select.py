"""A/B sweep of the selection kernel's ring feeders and geometry on one GPU (env knobs of csrc/selection.cu):
us/pick of FPS / k-center at N = 500k for the cp.async ring and the TMA feeders at several ring depths / warp counts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssdr_al_b200 import device as D

dev = torch.device("cuda", 0)
N = 500_000
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(kind, d, picks, n=N):
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    F = torch.randn((n, d), generator=g, device=dev, dtype=torch.float32)
    sel = torch.arange(n - 16, n, device=dev)
    fn = (lambda: D.fps(F, picks, 12345)) if kind == "fps" else (lambda: D.kcenter(F, sel, picks))
    fn()
    best = 1e9
    for _ in range(3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return 1e3 * best / picks


# fixed cost of a pick (grid-wide exchange + centre fetch): one 32-row unit per warp
for w in (4, 8, 16):
    os.environ["SSDR_SEL_FEED"], os.environ["SSDR_SEL_WARPS"] = "1", str(w)
    os.environ.pop("SSDR_SEL_STAGES", None)
    print("fixed cost: fps d=32 N=%d warps=%d %.2f us/pick" % (148 * w * 32, w, run("fps", 32, 3000, 148 * w * 32)), flush=True)
for kind, d, picks in (("fps", 32, 2000), ("kc", 32, 2000), ("fps", 256, 600), ("kc", 256, 600), ("fps", 64, 1000),
                       ("fps", 128, 800)):
    os.environ["SSDR_SEL_FEED"] = "0"
    os.environ.pop("SSDR_SEL_STAGES", None)
    os.environ.pop("SSDR_SEL_WARPS", None)
    print("%s d=%d feed=cp.async            %.2f us/pick" % (kind, d, run(kind, d, picks)), flush=True)
    os.environ["SSDR_SEL_FEED"] = "1"
    for st in (2, 3, 4):
        for w in (4, 6, 8, 12, 16):
            os.environ["SSDR_SEL_STAGES"], os.environ["SSDR_SEL_WARPS"] = str(st), str(w)
            try:
                print("%s d=%d feed=tma stages=%d warps=%2d %.2f us/pick" % (kind, d, st, w, run(kind, d, picks)), flush=True)
            except Exception as e:  # noqa: BLE001
                print("%s d=%d feed=tma stages=%d warps=%2d failed: %r" % (kind, d, st, w, e), flush=True)
