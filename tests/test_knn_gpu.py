"""GPU parity: KNN through the C ABI vs golden vectors (made by the reference nanoflann build) and the oracle.
Indices must be BIT-EXACT, including rows where equal fp32 distances make nanoflann's tree-visit order decide."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CLOUDS = ["uniform", "room", "quant", "dups", "tiny"]


@pytest.fixture(scope="module")
def NN():
    import ssdr_al_b200 as S
    return S.nearest_neighbors


def _room(rng, n, quant=None):
    face = rng.integers(0, 3, n)
    p = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
    p[face == 0, 2] = 0.0
    p[face == 1, 1] = 0.0
    p[face == 2, 0] = 0.0
    p += rng.normal(0, 0.005, p.shape)
    if quant:
        p = np.round(p / quant) * quant
    return p.astype(np.float32)


def _device_tree(pts):
    from ssdr_al_b200 import _lib
    pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts)
    cap = 3 * n + 64
    vind = np.zeros(n, np.uint32)
    nn = np.zeros(1, np.uint32)
    a = dict(left=np.zeros(cap, np.uint32), right=np.zeros(cap, np.uint32), child1=np.zeros(cap, np.int32),
             child2=np.zeros(cap, np.int32), divfeat=np.zeros(cap, np.int32), divlow=np.zeros(cap, np.float32),
             divhigh=np.zeros(cap, np.float32))
    _lib.check(_lib.lib().ssdr_knn_debug_tree(_lib.ptr(pts), n, _lib.ptr(vind), _lib.ptr(nn), _lib.ptr(a["left"]),
                                              _lib.ptr(a["right"]), _lib.ptr(a["child1"]), _lib.ptr(a["child2"]),
                                              _lib.ptr(a["divfeat"]), _lib.ptr(a["divlow"]), _lib.ptr(a["divhigh"])))
    return vind, {k: v[:nn[0]] for k, v in a.items()}


def _node_table(nodes):
    out = {}
    for i in range(len(nodes["left"])):
        key = (int(nodes["left"][i]), int(nodes["right"][i]))
        out[key] = None if nodes["child1"][i] < 0 else (int(nodes["divfeat"][i]), float(nodes["divlow"][i]),
                                                       float(nodes["divhigh"][i]))
    return out


@pytest.mark.parametrize("name", CLOUDS)
def test_device_tree_equals_sequential_tree(oracle, golden, name):
    p = golden.knn[name + "_pts"]
    vind, nodes = _device_tree(p)
    ovind, onodes = oracle.kdtree_export(p)
    assert np.array_equal(vind.astype(np.int64), ovind)
    assert _node_table(nodes) == _node_table(onodes)


def test_device_tree_degenerate_clouds(oracle):
    rng = np.random.default_rng(9)
    for p in (np.ones((700, 3), np.float32),
              np.stack([rng.random(900), np.zeros(900), np.zeros(900)], 1).astype(np.float32),
              rng.random((11, 3), dtype=np.float32), rng.random((1, 3), dtype=np.float32),
              rng.random((50_000, 3), dtype=np.float32)):
        vind, nodes = _device_tree(p)
        ovind, onodes = oracle.kdtree_export(p)
        assert np.array_equal(vind.astype(np.int64), ovind)
        assert _node_table(nodes) == _node_table(onodes)


@pytest.mark.parametrize("name", CLOUDS)
def test_knn_golden_self_k16(NN, golden, name):
    g = golden.knn
    p, want = g[name + "_pts"], g[name + "_k16"]
    got = NN.knn(p, p, want.shape[1], omp=True)
    assert got.dtype == np.int64 and got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", CLOUDS)
def test_knn_golden_external_queries(NN, golden, name):
    g = golden.knn
    p, q = g[name + "_pts"], g[name + "_q"]
    assert np.array_equal(NN.knn(p, q, 1), g[name + "_q_k1"])
    assert np.array_equal(NN.knn(p, q, 5), g[name + "_q_k5"])


def test_knn_golden_batch_and_upsampling(NN, golden):
    g = golden.knn
    bp = g["batch_pts"]
    assert np.array_equal(NN.knn_batch(bp, bp, 16, omp=True), g["batch_k16"])
    sub = bp[:, :300, :]  # non-contiguous slice, exactly like s3dis_dataset.py:166
    assert not sub.flags["C_CONTIGUOUS"]
    assert np.array_equal(NN.knn_batch(sub, bp, 1, omp=True), g["batch_sub_k1"])


def test_knn_k_larger_than_npts(NN, golden):
    g = golden.knn
    p = g["small_pts"]
    got = NN.knn(p, p[:1], 10)
    assert np.array_equal(got, g["small_k10"])


@pytest.mark.parametrize("K", [1, 3, 16, 20, 40])
def test_knn_vs_oracle_tie_free_and_tied(NN, oracle, K):
    rng = np.random.default_rng(K)
    clouds = {
        "uniform": rng.random((60_000, 3), dtype=np.float32) * np.float32(4.0) - np.float32(2.0),
        "room": _room(rng, 60_000),
        "quant": _room(rng, 30_000, quant=0.01),
        "dups": rng.random((4000, 3), dtype=np.float32)[rng.integers(0, 4000, 30_000)],
        "plane": np.concatenate([rng.random((20_000, 2), dtype=np.float32), np.zeros((20_000, 1), np.float32)], 1),
    }
    for name, p in clouds.items():
        want = oracle.knn(p, p, K, threads=8)
        got = NN.knn(p, p, K)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, "%s: %d rows differ, first %d: %s vs %s" % (name, bad.size, bad[0], got[bad[0]], want[bad[0]])


@pytest.mark.parametrize("K", [1, 5, 12, 16, 24])
def test_main_kernel_alone_is_right_on_tie_free_clouds(oracle, K):
    """The exact tie path re-resolves every flagged row, so it would hide a main kernel that duplicates or drops
    candidates (such rows hold equal distances and get flagged).  On clouds without ties only the boundary near-ties
    may be flagged: a few rows in 10^5, not a sizeable fraction -- and the result must still be the oracle's."""
    import torch
    from ssdr_al_b200 import device as D
    rng = np.random.default_rng(100 + K)
    for name, p in (("uniform", rng.random((2, 50_000, 3), dtype=np.float32) * np.float32(3.0)),
                    ("room", np.stack([_room(rng, 30_000), _room(rng, 30_000)]))):
        t = torch.from_numpy(p).cuda()
        got, st = D.knn_batch(t, t, K, want_stats=True)
        assert st["tie_rows"] <= 2e-3 * st["queries"], (name, K, st)
        want = oracle.knn_batch(p, p, K, threads=8)
        assert np.array_equal(got.cpu().numpy(), want), (name, K)


def test_knn_external_queries_outside_bbox(NN, oracle):
    rng = np.random.default_rng(77)
    p = _room(rng, 40_000)
    q = (rng.random((20_000, 3)) * np.array([12.0, 9.0, 6.0]) - 2.5).astype(np.float32)
    for K in (1, 16):
        assert np.array_equal(NN.knn(p, q, K), oracle.knn(p, q, K, threads=8))


def _ref_knn_batch(oracle):
    """The reference's own cpp_knn_batch_omp when oracle/_ref was built, else the C restatement."""
    if oracle.have_ref():
        oracle.set_omp_threads(6)
        return lambda p, q, k: oracle.ref_knn_batch(p, q, k, omp=True)
    return lambda p, q, k: oracle.knn_batch(p, q, k, threads=8)


def test_knn_batch_randla_pyramid_config2_full_size(NN, oracle):
    """BASELINE config 2 at its stated size (6 x 40960, five levels, k = 16 + 1-NN up-sampling): the loop of
    s3dis_dataset.py:164-177, every call compared with the reference."""
    rng = np.random.default_rng(1)
    B, N = 6, 40960
    xyz = (rng.uniform(-1, 1, (B, N, 3)) * np.array([2.0, 2.0, 1.5])).astype(np.float32)
    ref = _ref_knn_batch(oracle)
    for ratio in (4, 4, 4, 4, 2):
        neigh = NN.knn_batch(xyz, xyz, 16, omp=True)
        assert np.array_equal(neigh, ref(xyz, xyz, 16))
        sub = xyz[:, : xyz.shape[1] // ratio, :]
        up = NN.knn_batch(sub, xyz, 1, omp=True)
        assert np.array_equal(up, ref(np.ascontiguousarray(sub), xyz, 1))
        xyz = sub


@pytest.mark.parametrize("kind", ["uniform", "duplicated", "quantised"])
def test_knn_pyramid_single_call_equals_reference(oracle, kind):
    """ssdr_knn_pyramid_dev: the ten queries of the pyramid enqueued by ONE call without a host round trip (the trees
    of a level built at most once for its two queries, decided on the device).  Tie-free, duplicate-heavy (the
    loader's data_aug repeats points, s3dis_dataset.py:147-150) and quantised clouds; twice in a row on different
    data so that nothing stale survives between calls."""
    import torch
    from ssdr_al_b200 import _lib, device as dev
    ref = _ref_knn_batch(oracle)
    ratios = (4, 4, 4, 4, 2)
    for seed in (3, 4):
        rng = np.random.default_rng(seed)
        B, N = 6, 40960
        xyz = (rng.uniform(-1, 1, (B, N, 3)) * np.array([2.0, 2.0, 1.5])).astype(np.float32)
        if kind == "duplicated":
            for b in range(B):
                xyz[b, N // 2:] = xyz[b, rng.integers(0, N // 2, N - N // 2)]
                xyz[b] = xyz[b, rng.permutation(N)]
        elif kind == "quantised":
            xyz = (np.round(xyz * 64) / 64).astype(np.float32)
        neigh, up = dev.knn_pyramid(torch.from_numpy(xyz).cuda(), ratios, 16, check=True)
        assert int(_lib.lib().ssdr_knn_pyramid_launches()) > 0
        cur = xyz
        for l, ratio in enumerate(ratios):
            assert np.array_equal(neigh[l].cpu().numpy(), ref(cur, cur, 16)), (kind, seed, l)
            sub = np.ascontiguousarray(cur[:, : cur.shape[1] // ratio, :])
            assert np.array_equal(up[l].cpu().numpy(), ref(sub, cur, 1)), (kind, seed, l)
            cur = sub


def test_knn_pyramid_host_call_equals_the_ten_calls(NN, oracle):
    """nearest_neighbors.knn_pyramid (host arrays in and out, one C-ABI call) against the loop of
    s3dis_dataset.py:164-177 run through the reference; three calls in a row so that the branches' workspaces, the
    copies on the branch streams and the pinned result pool are all reused."""
    ref = _ref_knn_batch(oracle)
    ratios = (4, 4, 4, 4, 2)
    for seed, (B, N) in ((11, (6, 40960)), (12, (2, 4096)), (13, (6, 40960))):
        rng = np.random.default_rng(seed)
        xyz = (rng.uniform(-1, 1, (B, N, 3)) * np.array([2.0, 2.0, 1.5])).astype(np.float32)
        if seed == 13:  # duplicate-heavy: most rows take the tie path
            for b in range(B):
                xyz[b, N // 2:] = xyz[b, rng.integers(0, N // 2, N - N // 2)]
        neigh, up = NN.knn_pyramid(xyz, ratios, 16)
        cur = xyz
        for l, ratio in enumerate(ratios):
            assert neigh[l].dtype == np.int64 and neigh[l].shape == (B, cur.shape[1], 16)
            assert np.array_equal(neigh[l], ref(cur, cur, 16)), (seed, l)
            sub = np.ascontiguousarray(cur[:, : cur.shape[1] // ratio, :])
            assert np.array_equal(up[l], ref(sub, cur, 1)), (seed, l)
            cur = sub
    with pytest.raises(RuntimeError):
        NN.knn_pyramid(xyz[:, :32], (4, 4), 16)  # level 1 would hold 8 points, fewer than K


def test_knn_pyramid_repeated_calls_replay_and_follow_new_data(oracle):
    """The device pyramid replays a captured graph from the second identical call on; refilling the SAME device
    buffer with new points (what a loader does) must give the new cloud's rows, with the branches running side by
    side in the replay as well."""
    import torch
    from ssdr_al_b200 import device as dev
    ref = _ref_knn_batch(oracle)
    ratios = (4, 4, 4, 4, 2)
    B, N = 6, 40960
    buf = torch.empty((B, N, 3), dtype=torch.float32, device="cuda")
    neigh = up = None
    for seed in (21, 22, 23, 24):
        rng = np.random.default_rng(seed)
        xyz = (rng.uniform(-1, 1, (B, N, 3)) * np.array([2.0, 2.0, 1.5])).astype(np.float32)
        buf.copy_(torch.from_numpy(xyz))
        neigh, up = dev.knn_pyramid(buf, ratios, 16, neigh=neigh, up=up, check=True)
    cur = xyz
    for l, ratio in enumerate(ratios):
        assert np.array_equal(neigh[l].cpu().numpy(), ref(cur, cur, 16)), l
        sub = np.ascontiguousarray(cur[:, : cur.shape[1] // ratio, :])
        assert np.array_equal(up[l].cpu().numpy(), ref(sub, cur, 1)), l
        cur = sub


def test_knn_config1_full_size_properties(NN):
    """Config-1-sized cloud (~107k barycentres): self is the first neighbour, rows ascend in distance, and a random
    sample of rows equals brute force."""
    rng = np.random.default_rng(0)
    p = np.unique(np.round(_room(rng, 400_000) / 0.04).astype(np.int32), axis=0).astype(np.float32) * np.float32(0.04)
    p += rng.normal(0, 0.003, p.shape).astype(np.float32)
    idx = NN.knn(p, p, 16)
    assert (idx[:, 0] == np.arange(len(p))).all()
    d = ((p[idx] - p[:, None, :]) ** 2).sum(-1)
    assert (np.diff(d, axis=1) >= -1e-7).all()
    for r in rng.integers(0, len(p), 50):
        bd = ((p - p[r]) ** 2).sum(-1)
        assert set(np.argsort(bd, kind="stable")[:16].tolist()) == set(idx[r].tolist())


def test_knn_interface_contract(NN):
    p = np.random.default_rng(0).random((100, 3))
    assert NN.knn(p.astype(np.float64), p[:7], 4).shape == (7, 4)  # float64 is coerced like np.ascontiguousarray(., f32)
    with pytest.raises(RuntimeError, match="dim"):
        NN.knn(np.zeros((10, 2), np.float32), np.zeros((3, 2), np.float32), 2)
    with pytest.raises(NotImplementedError):
        NN.knn_batch_distance_pick(p[None], 10, 4)


def test_tree_reuse_between_calls_is_content_checked(NN, oracle):
    """The tie path keeps the last trees and reuses them when the next call brings the same support cloud (the
    pyramid's 1-NN up-sampling followed by the k=16 self query).  Same shape with different content -- another
    batch, or one moved point -- must rebuild; every answer equals the oracle."""
    rng = np.random.default_rng(31)

    def cloud(seed):  # quantised coordinates: ties in most rows, so the tie path always runs
        r = np.random.default_rng(seed)
        return (np.round(r.random((4, 3000, 3)) * 40) / 40).astype(np.float32)

    a, b = cloud(1), cloud(2)
    q = rng.random((4, 5000, 3)).astype(np.float32)
    assert np.array_equal(NN.knn_batch(a, q, 1), oracle.knn_batch(a, q, 1, threads=4))       # builds the trees of a
    assert np.array_equal(NN.knn_batch(a, a, 16), oracle.knn_batch(a, a, 16, threads=4))     # same cloud: reuse
    assert np.array_equal(NN.knn_batch(a, q, 8), oracle.knn_batch(a, q, 8, threads=4))       # reuse again
    assert np.array_equal(NN.knn_batch(b, b, 16), oracle.knn_batch(b, b, 16, threads=4))     # same shape, new content
    c = b.copy()
    c[2, 1234] += np.float32(0.5)                                                              # one point moved
    assert np.array_equal(NN.knn_batch(c, c, 16), oracle.knn_batch(c, c, 16, threads=4))
    assert np.array_equal(NN.knn_batch(b, b, 16), oracle.knn_batch(b, b, 16, threads=4))     # and back
    one = a[:1]                                                                                # other batch size
    assert np.array_equal(NN.knn_batch(one, one, 16), oracle.knn_batch(one, one, 16, threads=4))


def test_knn_batch_accepts_slices_views_and_other_dtypes(NN, oracle):
    """Batch items that are slices of a larger array are uploaded without a host-side packing copy; every other
    layout (reversed, transposed, float64, lists) takes the reference's np.ascontiguousarray route.  Same answers."""
    rng = np.random.default_rng(8)
    base = rng.random((3, 5000, 3)).astype(np.float32)
    want = oracle.knn_batch(np.ascontiguousarray(base[:, :1200]), np.ascontiguousarray(base[:, 100:4100]), 5, threads=4)
    pts, q = base[:, :1200], base[:, 100:4100]
    assert not pts.flags["C_CONTIGUOUS"]
    assert np.array_equal(NN.knn_batch(pts, q, 5), want)
    assert np.array_equal(NN.knn_batch(pts.astype(np.float64), q.tolist(), 5), want)
    rev = np.ascontiguousarray(base[:, :1200][:, ::-1])[:, ::-1]          # negative stride view of the same values
    assert np.array_equal(NN.knn_batch(rev, q, 5), want)
    t = np.ascontiguousarray(base.transpose(1, 0, 2)).transpose(1, 0, 2)   # item-interleaved memory
    assert np.array_equal(NN.knn_batch(t[:, :1200], t[:, 100:4100], 5), want)
    self_want = oracle.knn_batch(np.ascontiguousarray(pts), np.ascontiguousarray(pts), 16, threads=4)
    assert np.array_equal(NN.knn_batch(pts, pts, 16), self_want)
    assert np.array_equal(NN.knn_batch(base[1:2, :1200], base[1:2, :1200], 16), self_want[1:2])
