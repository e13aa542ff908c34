"""GPU parity: grid subsampling through the C ABI vs golden vectors (made by the reference) and the oracle.
Voxel membership, counts, barycentres, mean features and label votes are compared BIT-EXACTLY; rows are aligned by
voxel key because the product emits ascending-key order while the reference emits libstdc++ hash order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ssdr_al_b200 as S
    return S.grid_subsampling


def _room(rng, n):
    face = rng.integers(0, 3, n)
    p = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
    p[face == 0, 2] = 0.0
    p[face == 1, 1] = 0.0
    p[face == 2, 0] = 0.0
    p += rng.normal(0, 0.005, p.shape)
    return p.astype(np.float32)


def _by_key(oracle, pts, dl, ref_rows):
    """Re-order reference-order rows (tuple of arrays) into ascending voxel-key order."""
    o = oracle.grid_subsample(pts, None, None, dl, order="reference", with_keys=True)
    perm = np.argsort(o[3], kind="stable")
    assert o[0].tobytes() == ref_rows[0].tobytes()  # the fixture rows are in reference order
    return [None if r is None else r[perm] for r in ref_rows], o[3][perm], o[4][perm]


def test_grid_golden_full(G, oracle, golden):
    g = golden.grid
    (wp, wf, wc), wkeys, wcounts = _by_key(oracle, g["pts"], 0.1, (g["out_pts"], g["out_rgb"], g["out_lab"]))
    (p, f, c), keys, counts = G.compute(g["pts"], features=g["rgb"], classes=g["lab"], sampleDl=0.1, return_keys=True)
    assert p.dtype == np.float32 and f.dtype == np.float32 and c.dtype == np.int32 and c.shape == (len(p), 1)
    assert np.array_equal(keys, wkeys) and np.array_equal(counts, wcounts)
    assert p.tobytes() == wp.tobytes()
    assert f.tobytes() == wf.tobytes()
    assert np.array_equal(c, wc)


def test_grid_golden_points_only_negative(G, oracle, golden):
    g = golden.grid
    (wp,), _, _ = _by_key(oracle, g["pts"] - 3.0, 0.04, (g["neg_out_pts"],))
    p = G.compute(g["pts"] - 3.0, sampleDl=0.04)
    assert isinstance(p, np.ndarray) and p.tobytes() == wp.tobytes()


def test_grid_golden_label_collisions_two_columns(G, oracle, golden):
    g = golden.grid
    (wp, wc), _, _ = _by_key(oracle, g["pts"], 0.25, (g["lab2_out_pts"], g["lab2_out_lab"]))
    p, c = G.compute(g["pts"], classes=g["lab2"], sampleDl=0.25)
    assert p.tobytes() == wp.tobytes()
    assert c.shape == wc.shape and np.array_equal(c, wc)


@pytest.mark.parametrize("n,dl,fdim", [(200_000, 0.04, 3), (50_000, 0.5, 5), (30_000, 0.02, 1), (1000, 10.0, 3)])
def test_grid_vs_oracle(G, oracle, n, dl, fdim):
    rng = np.random.default_rng(n)
    pts = _room(rng, n)
    feats = rng.integers(0, 256, (n, fdim)).astype(np.uint8)
    lab = rng.integers(0, 13, n).astype(np.uint8)
    wp, wf, wc, wk, wn = oracle.grid_subsample(pts, feats, lab, dl, order="key", with_keys=True)
    (p, f, c), k, cnt = G.compute(pts, features=feats, classes=lab, sampleDl=dl, return_keys=True)
    assert np.array_equal(k, wk) and np.array_equal(cnt, wn)
    assert p.tobytes() == wp.tobytes() and f.tobytes() == wf.tobytes() and np.array_equal(c, wc)


def test_grid_heavy_voxel_and_many_labels(G, oracle):
    rng = np.random.default_rng(11)
    n = 40_000
    pts = (rng.random((n, 3)) * 0.05).astype(np.float32)  # everything in a handful of voxels
    pts[:100] += 5.0
    lab = rng.integers(-20, 30, n).astype(np.int32)  # 50 distinct labels -> inner map rehashes 13 -> 29 -> 59
    wp, _, wc = oracle.grid_subsample(pts, None, lab, 0.04, order="key")
    p, c = G.compute(pts, classes=lab, sampleDl=0.04)
    assert p.tobytes() == wp.tobytes() and np.array_equal(c, wc)


def test_grid_wide_extent_needs_more_than_32_key_bits(G, oracle):
    rng = np.random.default_rng(12)
    n = 100_000
    pts = (rng.random((n, 3)) * np.array([200.0, 200.0, 30.0])).astype(np.float32)
    pts[: n // 2] = (rng.random((n // 2, 3)) * 2.0).astype(np.float32)  # dense near the "scanner"
    wp, _, _, wk, wn = oracle.grid_subsample(pts, None, None, 0.06, order="key", with_keys=True)
    assert int(wk.max()) > 2 ** 32
    p, k, cnt = G.compute(pts, sampleDl=0.06, return_keys=True)
    assert np.array_equal(k, wk) and np.array_equal(cnt, wn) and p.tobytes() == wp.tobytes()


def test_grid_wrapper_contract(G):
    pts = np.random.default_rng(1).random((100, 3))
    assert isinstance(G.compute(pts, sampleDl=0.2), np.ndarray)
    r = G.compute(pts, features=np.ones((100, 2)), sampleDl=0.2)
    assert isinstance(r, tuple) and len(r) == 2 and r[1].shape[1] == 2
    r = G.compute(pts, classes=np.zeros((100, 2), np.int64), sampleDl=0.2)
    assert len(r) == 2 and r[1].shape[1] == 2 and r[1].dtype == np.int32
    with pytest.raises(RuntimeError, match="points.shape is not"):
        G.compute(np.zeros((10, 2)), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="features.shape is not"):
        G.compute(pts, features=np.ones((99, 2)), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="classes.shape is not"):
        G.compute(pts, classes=np.zeros(7, np.int32), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="Error parsing method"):
        G.compute(pts, sampleDl=0.1, method="nope")
    with pytest.raises(TypeError):
        G.compute(pts, np.ones((100, 3)))  # everything after `points` is keyword-only (format "O|$OOfsi")
    with pytest.raises(RuntimeError, match="^Error$"):
        G.compute(np.zeros((0, 3)), sampleDl=0.1)


def test_grid_config1_full_size_properties(G):
    """BASELINE config 1 size (1M pts): counts sum to N, keys strictly ascending, features/labels in range."""
    rng = np.random.default_rng(0)
    n = 1_000_000
    pts = _room(rng, n)
    pts -= pts.min(0)
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    lab = (pts[:, 0] * 1.7).astype(np.uint8) % 13
    (p, f, c), k, cnt = G.compute(pts, features=rgb, classes=lab, sampleDl=0.04, return_keys=True)
    assert cnt.sum() == n and (np.diff(k.astype(np.int64)) > 0).all()
    assert f.min() >= 0 and f.max() <= 255 and c.min() >= 0 and c.max() < 13
    # subsampling the barycentres at the same cell size can only merge, never split
    p2, k2, cnt2 = G.compute(p, sampleDl=0.04, return_keys=True)
    assert len(p2) <= len(p) and cnt2.sum() == len(p)


def test_grid_reference_row_order_golden(G, golden):
    """order="reference": rows come out in the reference's own libstdc++ hash-iteration order -- the fixture rows,
    produced by the unmodified reference, must be reproduced bit for bit INCLUDING their order."""
    g = golden.grid
    p, f, c = G.compute(g["pts"], features=g["rgb"], classes=g["lab"], sampleDl=0.1, order="reference")
    assert p.tobytes() == g["out_pts"].tobytes()
    assert f.tobytes() == g["out_rgb"].tobytes()
    assert np.array_equal(c, g["out_lab"])
    p = G.compute(g["pts"] - 3.0, sampleDl=0.04, order="reference")
    assert p.tobytes() == g["neg_out_pts"].tobytes()
    p, c = G.compute(g["pts"], classes=g["lab2"], sampleDl=0.25, order="reference")
    assert p.tobytes() == g["lab2_out_pts"].tobytes() and np.array_equal(c, g["lab2_out_lab"])


@pytest.mark.parametrize("n,dl", [(300_000, 0.04), (50_000, 0.3), (20, 0.5), (2_000_000, 0.05)])
def test_grid_reference_row_order_vs_oracle(G, oracle, n, dl):
    rng = np.random.default_rng(n + 1)
    pts = _room(rng, n)
    lab = rng.integers(0, 13, n).astype(np.uint8)
    wp, _, wc = oracle.grid_subsample(pts, None, lab, dl, order="reference")
    p, c = G.compute(pts, classes=lab, sampleDl=dl, order="reference")
    assert p.shape == wp.shape and p.tobytes() == wp.tobytes() and np.array_equal(c, wc)
