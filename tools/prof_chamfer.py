"""Chamfer adjacency timing: python tools/prof_chamfer.py [S]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ssdr_al_b200 as S_

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(9)
    sizes = rng.integers(60, 700, S)
    sps, cents = [], []
    for n in sizes:
        c = rng.random(3) * np.array([7.0, 5.0, 3.0])
        p = (c + rng.normal(0, 0.2, (int(n), 3)) * rng.choice([1.0, 0.05], 3)).astype(np.float32)
        sps.append(p); cents.append((p.min(0).astype(np.float64) + p.max(0)) / 2.0)
    cents = np.array(cents)
    S_.chamfer.create_cd(sps, cents)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); S_.chamfer.create_cd(sps, cents); ts.append(time.perf_counter() - t0)
    T = int(sizes.sum())
    print("S=%d T=%d create_cd %.2f ms  %.3g pair evals/s" % (S, T, 1e3 * np.median(ts), T * T / np.median(ts)))

if __name__ == "__main__":
    main()
