"""RandLA-style pyramid on SURFACE crops (what the real pipeline feeds: the 40960 nearest points of a random centre in
a room made of planes), device resident; compare with SSDR_KNN_PROBE=0.   python tools/prof_pyramid_surface.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ssdr_al_b200 import device as D


def room(rng, n):
    k = rng.integers(0, 8, n)
    p = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
    p[k == 0, 2] = 0.0
    p[k == 1, 2] = 3.0
    p[k == 2, 1] = 0.0
    p[k == 3, 1] = 5.0
    p[k == 4, 0] = 0.0
    p[k == 5, 0] = 7.0
    box = k >= 6
    p[box, 2] = 0.75
    p[box, 0] = 1.0 + (p[box, 0] % 2.0)
    p[box, 1] = 1.0 + (p[box, 1] % 1.2)
    p += rng.normal(0, 0.004, p.shape)
    return p.astype(np.float32)


def main():
    rng = np.random.default_rng(5)
    cloud = room(rng, 400_000)
    items = []
    for _ in range(6):
        c = cloud[rng.integers(0, len(cloud))]
        d = ((cloud - c) ** 2).sum(1)
        idx = np.argpartition(d, 40960)[:40960]
        items.append(cloud[rng.permutation(idx)])
    xyz0 = torch.from_numpy(np.stack(items)).cuda()

    def pyramid(stats=None):
        xyz = xyz0
        for ratio in (4, 4, 4, 4, 2):
            r = D.knn_batch(xyz, xyz, 16, want_stats=stats is not None)
            sub = xyz[:, : xyz.shape[1] // ratio, :].contiguous()
            r1 = D.knn_batch(sub, xyz, 1, want_stats=stats is not None)
            if stats is not None:
                stats.append(r[1])
                stats.append(r1[1])
            xyz = sub

    for _ in range(3):
        pyramid()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pyramid()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = []
    pyramid(st)
    print("probe=%s  pyramid %.3f ms (median of 10)" % (os.environ.get("SSDR_KNN_PROBE", "1"), float(np.median(ts))))
    print("  main ms:", [round(s["main_kernel_ms"], 3) for s in st])
    print("  grid ms:", [round(s["grid_build_ms"], 3) for s in st])
    print("  evals/q:", [int(s["dist_evals"] / max(1, s["queries"])) for s in st], " tie rows:", [int(s["tie_rows"]) for s in st])


if __name__ == "__main__":
    main()
