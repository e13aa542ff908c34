"""CPU: the C/numpy oracle restatement against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle (the reference itself ships no KATs)."""
import numpy as np
import pytest

CLOUDS = ["uniform", "room", "quant", "dups", "tiny"]


@pytest.mark.parametrize("name", CLOUDS)
def test_knn_self_k16(oracle, golden, name):
    g = golden.knn
    p = g[name + "_pts"]
    want = g[name + "_k16"]
    got = oracle.knn(p, p, want.shape[1])
    assert np.array_equal(got, want.astype(np.int64))


@pytest.mark.parametrize("name", CLOUDS)
def test_knn_external_queries(oracle, golden, name):
    g = golden.knn
    p, q = g[name + "_pts"], g[name + "_q"]
    assert np.array_equal(oracle.knn(p, q, 1, threads=4), g[name + "_q_k1"].astype(np.int64))
    assert np.array_equal(oracle.knn(p, q, 5), g[name + "_q_k5"].astype(np.int64))


def test_knn_batch(oracle, golden):
    g = golden.knn
    bp = g["batch_pts"]
    assert np.array_equal(oracle.knn_batch(bp, bp, 16, threads=3), g["batch_k16"].astype(np.int64))
    sub = np.ascontiguousarray(bp[:, :300, :])
    assert np.array_equal(oracle.knn_batch(sub, bp, 1), g["batch_sub_k1"].astype(np.int64))


def test_knn_k_larger_than_npts(oracle, golden):
    g = golden.knn
    p = g["small_pts"]
    got = oracle.knn(p, p[:1], 10)
    assert np.array_equal(got, g["small_k10"].astype(np.int64))
    assert (got[0, 7:] == 0).all()


def test_grid_full_reference_order(oracle, golden):
    g = golden.grid
    po, fo, co = oracle.grid_subsample(g["pts"], g["rgb"], g["lab"], 0.1)
    assert po.tobytes() == g["out_pts"].tobytes()  # bit-exact incl. row order
    assert fo.tobytes() == g["out_rgb"].tobytes()
    assert np.array_equal(co, g["out_lab"])
    assert co.shape[1] == 1 and co.dtype == np.int32


def test_grid_points_only_negative_coords(oracle, golden):
    g = golden.grid
    po, fo, co = oracle.grid_subsample(g["pts"] - 3.0, None, None, 0.04)
    assert fo is None and co is None
    assert po.tobytes() == g["neg_out_pts"].tobytes()


def test_grid_labels_colliding_mod13_two_columns(oracle, golden):
    g = golden.grid
    po, _, co = oracle.grid_subsample(g["pts"], None, g["lab2"], 0.25)
    assert po.tobytes() == g["lab2_out_pts"].tobytes()
    assert np.array_equal(co, g["lab2_out_lab"])


def test_grid_key_order_is_permutation_of_reference_order(oracle, golden):
    g = golden.grid
    a = oracle.grid_subsample(g["pts"], g["rgb"], g["lab"], 0.1, order="reference", with_keys=True)
    b = oracle.grid_subsample(g["pts"], g["rgb"], g["lab"], 0.1, order="key", with_keys=True)
    assert (np.diff(b[3].astype(np.int64)) > 0).all()
    perm = np.argsort(a[3], kind="stable")
    for x, y in zip(a, b):
        assert np.array_equal(x[perm], y)
    keys, _, _ = oracle.voxel_keys(g["pts"], 0.1)
    assert np.array_equal(np.unique(keys), b[3])
    assert np.array_equal(np.unique(keys, return_counts=True)[1], b[4])


@pytest.mark.parametrize("tag", ["d32", "d256", "d129_f64", "d3", "d20"])
def test_fps(oracle, golden, tag):
    g = golden.fps
    F, want = g[tag + "_F"], g[tag + "_picks"]
    assert np.array_equal(oracle.fps(F, len(want), int(want[0])), want)
    assert np.array_equal(oracle.fps_numpy(F, len(want), int(want[0])), want)


@pytest.mark.parametrize("D", [3, 7, 8, 9, 32, 100, 128, 129, 136, 256, 300, 1000])
def test_pairwise_sum_order_matches_numpy(oracle, D):
    rng = np.random.default_rng(D)
    F = rng.standard_normal((257, D)).astype(np.float32)
    assert np.array_equal(np.sum((F - F[5]) ** 2, axis=-1), oracle.rowdist(F, 5))
    F64 = F.astype(np.float64)
    assert np.array_equal(np.sum((F64 - F64[5]) ** 2, axis=-1), oracle.rowdist(F64, 5))


@pytest.mark.parametrize("tag", ["d129_f64", "d32_f32", "d256_f64"])
def test_kcenter(oracle, golden, tag):
    g = golden.kcenter
    X, sel, want = g[tag + "_X"], g[tag + "_sel"], g[tag + "_picks"]
    assert np.array_equal(oracle.kcenter(X, sel, len(want)), want)


def test_hash_order_small_cases(oracle):
    # all keys in distinct buckets of the 13-bucket table -> reverse insertion order
    assert list(oracle.hash_order([3, 7, 1, 12])) == [3, 2, 1, 0]
    # 1 and 14 share bucket 1: the later one goes to the FRONT of that bucket's chain, bucket keeps its place
    assert list(oracle.hash_order([1, 5, 14])) == [1, 2, 0]


@pytest.mark.parametrize("tag", ["small", "room"])
def test_chamfer_restatement_equals_reference_golden(oracle, golden, tag):
    """oracle.create_cd (brute force, numpy) vs the matrices the unmodified reference produced with its KD trees."""
    g = golden.chamfer
    pts, off = g[tag + "_points"], g[tag + "_offsets"]
    sps = [pts[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    got = oracle.create_cd(sps, g[tag + "_centroids"])
    np.testing.assert_allclose(got, g[tag + "_cd"], rtol=1e-12, atol=0)
    assert np.mean(got == g[tag + "_cd"]) > 0.99


@pytest.mark.parametrize("tag", ["small", "room"])
def test_superpoint_fps_restatement_equals_reference_golden(oracle, golden, tag):
    g = golden.chamfer
    pts, off = g[tag + "_points"], g[tag + "_offsets"]
    sps = [pts[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    want = g[tag + "_fps_picks"]
    got = oracle.farthest_superpoint_sample(sps, g[tag + "_centroids"], len(want), int(g[tag + "_fps_trigger"]))
    assert np.array_equal(got, want)
