"""Prototype of the data-parallel (epoch by epoch, sort based) emulation of libstdc++ unordered_map iteration order
that grid.cu uses for SSDR_GRID_ORDER_REFERENCE; checked against the sequential emulation in the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O

BUCKETS = [13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933, 351061, 712697, 1447153,
           2938679, 5967347, 12117689, 24607243, 49969847, 101473717, 206062531, 418451333, 849749479, 1725587117,
           3504151727]


def hash_order_parallel(keys):
    keys = np.asarray(keys, np.uint64)
    M = len(keys)
    cur = np.zeros(0, np.int64)  # node ids (insertion ranks) in current list order
    for nb in BUCKETS:
        prev = len(cur)
        if prev >= M:
            break
        hi = min(nb, M)
        seq = np.concatenate([cur, np.arange(prev, hi)])  # S: re-insertion of the list, then the new nodes
        s = np.arange(hi)
        b = (keys[seq] % np.uint64(nb)).astype(np.int64)
        cmin = np.full(nb, hi, np.int64)
        np.minimum.at(cmin, b, s)
        comp = (hi - 1 - cmin[b]) * hi + (hi - 1 - s)  # bucket creation descending, then sequence index descending
        cur = seq[np.argsort(comp, kind="stable")]
    return cur


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    ok = True
    for M in (1, 5, 13, 14, 29, 30, 100, 1000, 5000, 200000):
        for trial in range(3):
            keys = rng.choice(np.arange(0, max(4 * M, 50), dtype=np.uint64) * np.uint64(rng.integers(1, 5)), M, replace=False)
            want = O.hash_order(keys)
            got = hash_order_parallel(keys)
            ok &= bool(np.array_equal(want, got))
    print("closed form == sequential emulation:", ok)
