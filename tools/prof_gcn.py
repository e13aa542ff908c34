"""Timing of the adjacency / propagation kernels (csrc/gcn.cu) at a few sizes: python tools/prof_gcn.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssdr_al_b200 as S
G = S.fps_gcn
rng = np.random.default_rng(1)
for n_sp, d in ((4096, 32), (8192, 32), (16384, 32), (16384, 128)):
    per = 32
    perm = rng.permutation(n_sp)
    rooms = []
    for r0 in range(0, n_sp, per):
        cd = rng.random((per, per)); cd = cd + cd.T; np.fill_diagonal(cd, 0.0)
        rooms.append((perm[r0:r0 + per].tolist(), rng.random((per, 3)) * 8.0, cd))
    v = rng.standard_normal((n_sp, d))
    a = G.adjacency_from_rooms(n_sp, rooms)
    t0 = time.perf_counter(); a2 = G.adjacency_from_rooms(n_sp, rooms); t_adj = (time.perf_counter() - t0) * 1e3; a2.close()
    G.propagate(a, v, 1, 0)
    ts = {}
    for g in (1, 3):
        t0 = time.perf_counter(); G.propagate(a, v, g, 0); ts[g] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); G.propagate(a, v, 1, 8); t_top = (time.perf_counter() - t0) * 1e3
    per_prod = (ts[3] - ts[1]) / 2
    print("N=%d D=%d: adjacency %.2f ms, propagate g=1 %.2f ms, g=3 %.2f ms -> %.3f ms per product = %.0f GB/s of the matrix, "
          "top-8 mask + 1 product %.2f ms" % (n_sp, d, t_adj, ts[1], ts[3], per_prod, 8.0 * n_sp * n_sp / per_prod / 1e6, t_top), flush=True)
    a.close()
