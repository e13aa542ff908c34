"""Small workload touching every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ssdr_al_b200 as S


def main():
    rng = np.random.default_rng(0)
    # KNN: quantised clouds -> most rows tied -> tree build (top levels need > 4096 points) + exact replay
    big = (np.round(rng.random((2, 9000, 3)) * 30) / 30).astype(np.float32)
    S.nearest_neighbors.knn_batch(big, big, 16)
    small = (np.round(rng.random((3, 700, 3)) * 12) / 12).astype(np.float32)
    q = rng.random((3, 1500, 3)).astype(np.float32)
    S.nearest_neighbors.knn_batch(small, q, 1)       # builds the trees
    S.nearest_neighbors.knn_batch(small, small, 16)  # reuses them
    S.nearest_neighbors.knn(rng.random((50000, 3)).astype(np.float32), rng.random((2000, 3)).astype(np.float32), 8)
    # grid subsampling, both row orders, uint8 inputs
    p = (rng.random((60000, 3)) * [7, 5, 3]).astype(np.float32)
    rgb = rng.integers(0, 256, (60000, 3)).astype(np.uint8)
    lab = rng.integers(0, 13, 60000).astype(np.uint8)
    S.grid_subsampling.compute(p, features=rgb, classes=lab, sampleDl=0.2)
    S.grid_subsampling.compute(p, features=rgb, classes=lab, sampleDl=0.2, order="reference")
    # selection
    F = rng.standard_normal((20000, 32)).astype(np.float32)
    S.selection.fps(F, 40, 7)
    S.selection.kcenter(F.astype(np.float64), np.arange(5), 20)
    S.selection.fps(rng.standard_normal((3000, 100)).astype(np.float32), 20, 1)
    # chamfer adjacency
    sps = [rng.random((int(n), 3)).astype(np.float32) + i for i, n in enumerate((5, 130, 300, 1100))]
    S.chamfer.create_cd(sps, np.array([p.mean(0) for p in sps], np.float64))
    S.chamfer.farthest_superpoint_sample(sps, np.array([p.mean(0) for p in sps], np.float64), 3, 1)
    print("sanitize_driver done", flush=True)


if __name__ == "__main__":
    main()
