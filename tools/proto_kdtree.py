"""Prototype (numpy, flat data-parallel passes) of the level-synchronous nanoflann-identical tree build that
ssdr_al_b200/csrc/kdtree.cuh implements on the device.  Checked against the oracle's sequential build."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O

f32 = np.float32
LEAF = 10


def build(pts):
    pts = np.ascontiguousarray(pts, f32)
    N = len(pts)
    vind = np.arange(N, dtype=np.int64)
    node_of = np.zeros(N, np.int64)  # node id per position
    # node arrays
    L = [0]; R = [N]
    lo = [pts.min(0)]; hi = [pts.max(0)]           # loose bbox (root: data bbox)
    tlo = [None]; thi = [None]
    c1 = [-1]; c2 = [-1]; feat = [0]; parent = [-1]
    level_nodes = [0]
    while level_nodes:
        # P1: tight bbox of this level's nodes
        for n in level_nodes:
            seg = pts[vind[L[n]:R[n]]]
            tlo[n] = seg.min(0); thi[n] = seg.max(0)
        active = [n for n in level_nodes if R[n] - L[n] > LEAF]
        if not active:
            break
        # P2: decisions
        cutfeat = {}; cutval = {}
        for n in active:
            span = hi[n] - lo[n]
            max_span = span.max()
            best = -1; spread_best = f32(-1)
            for d in range(3):
                if span[d] > f32(f32(1) - f32(0.00001)) * max_span:
                    spread = thi[n][d] - tlo[n][d]
                    if spread > spread_best:
                        best = d; spread_best = spread
            if best < 0: best = 0
            cf = best
            sv = f32(f32(lo[n][cf] + hi[n][cf]) / f32(2))
            mn, mx = tlo[n][cf], thi[n][cf]
            cv = mn if sv < mn else (mx if sv > mx else sv)
            cutfeat[n] = cf; cutval[n] = cv
        # per position values
        pos = np.arange(N)
        is_act = np.zeros(N, bool)
        cf_pos = np.zeros(N, np.int64); cv_pos = np.zeros(N, f32); l_pos = np.zeros(N, np.int64); r_pos = np.zeros(N, np.int64)
        for n in active:
            s = slice(L[n], R[n]); is_act[s] = True; cf_pos[s] = cutfeat[n]; cv_pos[s] = cutval[n]; l_pos[s] = L[n]; r_pos[s] = R[n]
        lim = {}
        start_pos = l_pos.copy()
        for pass_id in (0, 1):
            val = pts[vind, cf_pos]
            sat = (val < cv_pos) if pass_id == 0 else (val <= cv_pos)
            inrange = is_act & (pos >= start_pos)
            satf = (sat & inrange).astype(np.int64); failf = ((~sat) & inrange).astype(np.int64)
            csat = np.cumsum(satf); cfail = np.cumsum(failf)            # inclusive prefix
            ex = lambda c, i: c[i - 1] if i > 0 else 0
            Lpos = np.full(N, -1, np.int64); Rpos = np.full(N, -1, np.int64)
            # per node totals
            tot_sat = {}
            for n in active:
                s0 = start_pos[L[n]]
                tot_sat[n] = csat[R[n] - 1] - ex(csat, s0)
            # flat pass
            for i in range(N):
                if not inrange[i]: continue
                n = node_of[i]; s0 = start_pos[i]
                limn = s0 + tot_sat[n]
                if not sat[i]:
                    k = cfail[i] - failf[i] - ex(cfail, s0)        # exclusive rank among failing
                    if i < limn: Lpos[L[n] + k] = i
                else:
                    k = tot_sat[n] - (csat[i] - ex(csat, s0))      # exclusive suffix rank among satisfying
                    if i >= limn: Rpos[L[n] + k] = i
            for n in active:
                k = 0
                while L[n] + k < R[n] and Lpos[L[n] + k] >= 0:
                    a, b = Lpos[L[n] + k], Rpos[L[n] + k]
                    assert b >= 0 and a < b
                    vind[a], vind[b] = vind[b], vind[a]
                    k += 1
                lim[(n, pass_id)] = start_pos[L[n]] + tot_sat[n]
            if pass_id == 0:
                for n in active:
                    start_pos[L[n]:R[n]] = lim[(n, 0)]
        nxt = []
        for n in active:
            count = R[n] - L[n]
            lim1 = lim[(n, 0)] - L[n]; lim2 = lim[(n, 1)] - L[n]
            idx = lim1 if lim1 > count // 2 else (lim2 if lim2 < count // 2 else count // 2)
            cf, cv = cutfeat[n], cutval[n]
            a = len(L); L.append(L[n]); R.append(L[n] + idx); l_ = lo[n].copy(); h_ = hi[n].copy(); h_[cf] = cv
            lo.append(l_); hi.append(h_); tlo.append(None); thi.append(None); c1.append(-1); c2.append(-1); feat.append(0)
            b = len(L); L.append(L[n] + idx); R.append(R[n]); l_ = lo[n].copy(); h_ = hi[n].copy(); l_[cf] = cv
            lo.append(l_); hi.append(h_); tlo.append(None); thi.append(None); c1.append(-1); c2.append(-1); feat.append(0)
            c1[n] = a; c2[n] = b; feat[n] = cf
            node_of[L[a]:R[a]] = a; node_of[L[b]:R[b]] = b
            nxt += [a, b]
        level_nodes = nxt
    nodes = {}
    for n in range(len(L)):
        if c1[n] >= 0:
            nodes[(L[n], R[n])] = (feat[n], thi[c1[n]][feat[n]], tlo[c2[n]][feat[n]])
        else:
            nodes[(L[n], R[n])] = None
    return vind, nodes


def check(pts, name):
    vind, nodes = build(pts)
    ov, on = O.kdtree_export(pts)
    ok_v = np.array_equal(vind, ov)
    onodes = {}
    for i in range(len(on["left"])):
        key = (int(on["left"][i]), int(on["right"][i]))
        onodes[key] = None if on["child1"][i] < 0 else (int(on["divfeat"][i]), on["divlow"][i], on["divhigh"][i])
    ok_n = set(nodes.keys()) == set(onodes.keys()) and all(
        (nodes[k] is None and onodes[k] is None) or (nodes[k] is not None and onodes[k] is not None and
         nodes[k][0] == onodes[k][0] and nodes[k][1] == onodes[k][1] and nodes[k][2] == onodes[k][2]) for k in nodes)
    print(name, "vind", ok_v, "nodes", ok_n, len(nodes))
    return ok_v and ok_n


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    ok = True
    ok &= check(rng.random((2000, 3), dtype=f32), "uniform")
    ok &= check((np.round(rng.random((2000, 3)) * [30, 30, 2]) / 10).astype(f32), "quantised")
    ok &= check(rng.random((100, 3), dtype=f32)[rng.integers(0, 100, 2000)], "dups")
    ok &= check(np.ones((500, 3), f32), "identical")
    ok &= check(np.stack([rng.random(700), np.zeros(700), np.zeros(700)], 1).astype(f32), "collinear")
    ok &= check(rng.random((11, 3), dtype=f32), "n11")
    ok &= check(rng.random((10, 3), dtype=f32), "n10")
    print("ALL OK" if ok else "MISMATCH")
