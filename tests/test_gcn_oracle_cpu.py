"""CPU: the numpy restatement of fps_adj_all / GCN_FPS_sampling's matrix part (oracle.gcn_adjacency, oracle.gcn_propagate)
against outputs of the reference itself (tests/golden/gcn.npz, made by tests/golden/make_golden_gcn.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixture():
    g = np.load(os.path.join(HERE, "golden", "gcn.npz"))
    rooms_xyz = {str(n): g[str(n) + "_xyz"] for n in g["room_names"]}
    comps = {}
    for n in rooms_xyz:
        ends = np.cumsum(g[n + "_sizes"])
        comps[n] = [np.arange(e - s, e) for s, e in zip(g[n + "_sizes"], ends)]
    unl = [{"cloud_name": str(c), "sp_idx": int(s)} for c, s in zip(g["unlabeled_cloud"], g["unlabeled_sp"])]
    lab = [{"cloud_name": str(c), "sp_idx": int(s)} for c, s in zip(g["labeled_cloud"], g["labeled_sp"])]
    return g, rooms_xyz, comps, unl, lab


def rooms_of(rooms_xyz, comps, unl, lab, create_cd):
    """The per-room quantities fps_adj_all holds after fps_gcn_cpu.py:91-92 (rows: unlabeled first, then labeled)."""
    members, names = {}, []
    for i, r in enumerate(unl + lab):
        if r["cloud_name"] not in members:
            members[r["cloud_name"]] = []
            names.append(r["cloud_name"])
        members[r["cloud_name"]].append((r["sp_idx"], i))
    out = []
    for n in names:
        xyz = rooms_xyz[n]
        centre = np.zeros([len(members[n]), 3])
        sps = []
        for j, (sp, _) in enumerate(members[n]):
            p = xyz[comps[n][sp]]
            for d in range(3):
                centre[j, d] = (np.min(p[:, d]) + np.max(p[:, d])) / 2.0
            sps.append(p)
        out.append(([i for _, i in members[n]], centre, create_cd(sps, centre)))
    return out


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def test_adjacency_restatement_equals_the_reference(oracle):
    g, rooms_xyz, comps, unl, lab = load_fixture()
    rooms = rooms_of(rooms_xyz, comps, unl, lab, oracle.create_cd)
    adj = oracle.gcn_adjacency(len(unl) + len(lab), rooms)
    # the chamfer blocks come from a brute-force restatement of the KD-tree queries: equal up to the trees' own rounding
    np.testing.assert_allclose(adj, g["adj"], rtol=1e-12, atol=0)
    assert np.array_equal(adj == 0, g["adj"] == 0)  # superpoints of different rooms: exp(-2e10) == 0 exactly


@pytest.mark.parametrize("g_top", [(1, 0), (2, 0), (1, 4)])
def test_propagation_restatement_equals_the_reference(oracle, g_top):
    g, _, _, unl, _ = load_fixture()
    gn, top = g_top
    v = np.concatenate([g["unlabeled_features"], g["labeled_features"]])
    got = oracle.gcn_propagate(g["adj"], v, gn, top)[:len(unl)]
    np.testing.assert_allclose(got, g["combo_g%d_top%d" % (gn, top)], rtol=1e-13, atol=1e-15)
