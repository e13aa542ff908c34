"""Phase timing of the device tree build (diagnostic): python tools/prof_tree.py B N"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ssdr_al_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40960
rng = np.random.default_rng(1)
pts = (rng.uniform(-1, 1, (B, N, 3)) * np.array([2.0, 2.0, 1.5])).astype(np.float32)
m = np.zeros(16, np.uint64)
_lib.check(_lib.lib().ssdr_knn_debug_build_timing(_lib.ptr(pts), B, N, _lib.ptr(m)))
t0 = int(m[0])
names = ["start", "roots"] + ["top%d" % k for k in range(8)] + ["top_done", "cta0_done", "all_done"]
print("B=%d N=%d" % (B, N), " ".join("%s=%.1fus" % (n, (int(v) - t0) / 1e3) for n, v in zip(names, m) if int(v)))
