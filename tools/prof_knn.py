"""Driver for ncu: N pyramids of bench.py's workload through the device-resident C ABI (no stats, no timing), so the
launch sequence is only this library's kernels plus the copy-engine packing of the levels.
    python tools/prof_knn.py [pyramids] [fused|calls]     fused = one ssdr_knn_pyramid_dev call per pyramid (default)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ssdr_al_b200 import device as D


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    fused = (sys.argv[2] if len(sys.argv) > 2 else "fused") == "fused"
    xyz0 = torch.from_numpy(bench.make_clouds(1)).cuda()
    for _ in range(reps):
        if fused:
            D.knn_pyramid(xyz0, bench.RATIOS, bench.K)
            continue
        xyz = xyz0
        for ratio in bench.RATIOS:
            D.knn_batch(xyz, xyz, bench.K)
            sub = xyz[:, : xyz.shape[1] // ratio, :].contiguous()
            D.knn_batch(sub, xyz, 1)
            xyz = sub
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
