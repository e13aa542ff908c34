"""Wall clock of the one-call host pyramid (nearest_neighbors.knn_pyramid), with the library's own phase trace
(SSDR_TRACE=1): python tools/prof_pyramid_host.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SSDR_TRACE"] = "1"
import numpy as np
import bench
import ssdr_al_b200 as S
from ssdr_al_b200 import _lib
NN = S.nearest_neighbors
xyz = bench.make_clouds(1)
pin = _lib.pinned_empty(xyz.shape, np.float32); pin[...] = xyz
for name, src in (("pageable", xyz), ("pinned", pin)):
    for _ in range(3):
        NN.knn_pyramid(src, bench.RATIOS, bench.K)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); r = NN.knn_pyramid(src, bench.RATIOS, bench.K); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "knn_pyramid wall ms: median %.3f min %.3f" % (np.median(ts), min(ts)), flush=True)
