"""Single large cloud (configs 1 and 3): k=16 self-KNN, device resident, with the stage breakdown.
    python tools/prof_knn_large.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ssdr_al_b200 import device as D


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100_000, 1_000_000, 4_000_000]
    rng = np.random.default_rng(0)
    for n in sizes:
        for name in ("volume", "surface"):
            p = rng.random((1, n, 3), dtype=np.float32) * np.array([20, 15, 6], np.float32)
            if name == "surface":
                p[0, : n // 2, 2] = 0.0
                p[0, n // 2:, 1] = 0.0
            a = torch.from_numpy(p).cuda()
            b = (a * 1.25).contiguous()  # alternate two clouds: no tree reuse between the timed calls
            D.knn_batch(a, a, 16)
            D.knn_batch(b, b, 16)
            best, st = 1e9, None
            for rep in range(4):
                c = b if rep % 2 else a
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _, s = D.knn_batch(c, c, 16, want_stats=True)
                e1.record()
                torch.cuda.synchronize()
                if e0.elapsed_time(e1) < best:
                    best, st = e0.elapsed_time(e1), s
            print("N=%-8d %-7s %.3f ms  %.1f M q/s | grid %.3f main %.3f tie %.3f (tree %.3f) rows %d evals/q %.0f" % (
                n, name, best, n / best / 1e3, st["grid_build_ms"], st["main_kernel_ms"], st["tie_path_ms"],
                st["tree_build_ms"], st["tie_rows"], st["dist_evals"] / n), flush=True)
            del a, b


if __name__ == "__main__":
    main()
