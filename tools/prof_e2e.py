"""Wall-clock per call of the host-API pyramid (what bench.py's e2e sums), next to the raw pinned PCIe copy rates.
    python tools/prof_e2e.py [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ssdr_al_b200 as S
from ssdr_al_b200 import _lib


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(0)
    B, N, K = 6, 40960, 16
    clouds = _lib.pinned_empty((B, N, 3), np.float32)
    clouds[...] = rng.random((B, N, 3), dtype=np.float32)
    NN = S.nearest_neighbors
    # raw copies
    d = torch.empty(32 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    h = torch.empty(32 * 1024 * 1024, dtype=torch.uint8).pin_memory()
    for nbytes in (1 << 20, 4 << 20, 32 << 20):
        for name, fn in (("d2h", lambda: h[:nbytes].copy_(d[:nbytes], non_blocking=True)),
                         ("h2d", lambda: d[:nbytes].copy_(h[:nbytes], non_blocking=True))):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 10
            print("raw pinned %s %5.1f MB: %.3f ms = %.1f GB/s" % (name, nbytes / 1e6, dt * 1e3, nbytes / dt / 1e9))
    times = {}
    for rep in range(reps + 3):
        xyz = clouds
        for lvl, ratio in enumerate((4, 4, 4, 4, 2)):
            t0 = time.perf_counter()
            idx = NN.knn_batch(xyz, xyz, K, omp=True)
            t1 = time.perf_counter()
            sub = xyz[:, : xyz.shape[1] // ratio, :]
            up = NN.knn_batch(sub, xyz, 1, omp=True)
            t2 = time.perf_counter()
            if rep >= 3:
                times.setdefault((lvl, "k16", xyz.shape[1], idx.nbytes), []).append(t1 - t0)
                times.setdefault((lvl, "k1", xyz.shape[1], up.nbytes), []).append(t2 - t1)
            xyz = sub
    tot = 0.0
    for key, v in times.items():
        med = float(np.median(v))
        tot += med
        print("level %d %-3s N=%-6d out %8.2f MB: %.3f ms" % (key[0], key[1], key[2], key[3] / 1e6, med * 1e3))
    print("sum of medians: %.3f ms per pyramid" % (tot * 1e3))


if __name__ == "__main__":
    main()
