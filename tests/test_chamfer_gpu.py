"""GPU parity: chamfer adjacency (create_cd, fps_gcn_cpu.py:12-38) through the C ABI vs golden matrices made by the
unmodified reference (sklearn KDTree) and vs the oracle restatement.  Distances are float64; the kernel reproduces the
reference's arithmetic order, so the comparison is exact up to the KD tree's own bound rounding: rtol 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _split(g, tag):
    pts, off = g[tag + "_points"], g[tag + "_offsets"]
    return [pts[off[i]:off[i + 1]] for i in range(len(off) - 1)], g[tag + "_centroids"], g[tag + "_cd"]


@pytest.mark.parametrize("tag", ["small", "room"])
def test_chamfer_golden(golden, tag):
    import ssdr_al_b200 as S
    sps, cents, want = _split(golden.chamfer, tag)
    got = S.chamfer.create_cd(sps, cents)
    assert got.shape == want.shape and got.dtype == np.float64
    assert (np.diag(got) == 0).all() and np.array_equal(got, got.T)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    assert np.mean(got == want) > 0.99  # in practice bit-identical; the tolerance only covers KD-tree bound rounding


def test_chamfer_vs_oracle_many_sizes(oracle):
    import ssdr_al_b200 as S
    rng = np.random.default_rng(3)
    sizes = [1, 3, 8, 57, 130, 255, 256, 257, 1023, 1024, 1025, 2049, 5000]
    sps, cents = [], []
    for n in sizes:
        c = rng.random(3) * 6
        p = (c + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
        sps.append(p)
        cents.append((p.min(0).astype(np.float64) + p.max(0)) / 2.0)
    got = S.chamfer.create_cd(sps, np.array(cents))
    want = oracle.create_cd(sps, np.array(cents))
    assert np.array_equal(got, want)  # same arithmetic order as the numpy restatement: bit-identical


def test_chamfer_interface(oracle):
    import ssdr_al_b200 as S
    assert S.chamfer.create_cd([], np.zeros((0, 3))).shape == (0, 0)
    one = S.chamfer.create_cd([np.ones((5, 3), np.float32)], np.zeros((1, 3)))
    assert one.shape == (1, 1) and one[0, 0] == 0
    with pytest.raises(RuntimeError, match="empty superpoint"):
        S.chamfer.create_cd([np.ones((5, 3)), np.zeros((0, 3))], np.zeros((2, 3)))
    # float64 inputs and list centroids
    rng = np.random.default_rng(0)
    sps = [rng.random((40, 3)), rng.random((60, 3)) + 2]
    cents = [list(s.mean(0)) for s in sps]
    assert np.array_equal(S.chamfer.create_cd(sps, cents), oracle.create_cd(sps, cents))


@pytest.mark.parametrize("tag", ["small", "room"])
def test_superpoint_fps_golden(golden, tag):
    """farthest_superpoint_sample (sampler2.py:49-80) vs the picks of the reference's own function."""
    import ssdr_al_b200 as S
    g = golden.chamfer
    sps, cents, _ = _split(g, tag)
    want = g[tag + "_fps_picks"]
    got = S.chamfer.farthest_superpoint_sample(sps, cents, len(want), int(g[tag + "_fps_trigger"]))
    assert got.dtype == np.int32 and np.array_equal(got, want)


def test_superpoint_fps_vs_oracle(oracle):
    import ssdr_al_b200 as S
    rng = np.random.default_rng(11)
    sps, cents = [], []
    for n in [int(v) for v in rng.integers(3, 500, 60)] + [1500]:
        c = rng.random(3) * 8
        p = (c + rng.normal(0, 0.3, (n, 3)) * rng.choice([1.0, 0.05], 3)).astype(np.float32)
        sps.append(p)
        cents.append((p.min(0).astype(np.float64) + p.max(0)) / 2.0)
    cents = np.array(cents)
    assert np.array_equal(S.chamfer.farthest_superpoint_sample(sps, cents, 25, 7),
                          oracle.farthest_superpoint_sample(sps, cents, 25, 7))
    assert np.array_equal(S.chamfer.farthest_superpoint_sample(sps, cents, 1, 5), np.array([5], np.int32))


def test_superpoint_fps_float32_centroids(oracle):
    """The real caller passes float32 centroids (ply coordinates, sampler2.py:563-577): the squared centroid distance is
    then a float32 quantity (np.sum((cents - cur) ** 2, -1)) that is only widened by np.add with the float64 chamfer
    row.  Near-ties between candidates are decided by that rounding, so the dtype has to be kept."""
    import ssdr_al_b200 as S
    rng = np.random.default_rng(21)
    sps, cents = [], []
    for n in [int(v) for v in rng.integers(5, 300, 80)]:
        c = rng.random(3) * 30
        p = (c + rng.normal(0, 0.25, (n, 3))).astype(np.float32)
        sps.append(p)
        cents.append(((p.min(0) + p.max(0)) / np.float32(2)).astype(np.float32))
    cents32 = np.array(cents, dtype=np.float32)
    want32 = oracle.farthest_superpoint_sample(sps, cents32, 40, 3)
    assert np.array_equal(S.chamfer.farthest_superpoint_sample(sps, cents32, 40, 3), want32)
    assert np.array_equal(S.chamfer.farthest_superpoint_sample(sps, list(cents32), 40, 3), want32)  # list of f32 rows
    cents64 = cents32.astype(np.float64)
    assert np.array_equal(S.chamfer.farthest_superpoint_sample(sps, cents64, 40, 3),
                          oracle.farthest_superpoint_sample(sps, cents64, 40, 3))


def test_chamfer_large_superpoint(oracle):
    """A floor or a wall can hold far more points than the shared leaf table of the pairwise mean covers (57 values per
    leaf at least, 1024 leaves): such clouds go through the windowed table; the reference has no size limit either."""
    import ssdr_al_b200 as S
    rng = np.random.default_rng(5)
    big = (rng.random((70_001, 3)) * np.array([7.0, 5.0, 0.02])).astype(np.float32)
    sps = [big] + [(rng.random(3) * 5 + rng.normal(0, 0.2, (n, 3))).astype(np.float32) for n in (40, 333, 1500)]
    cents = np.array([(p.min(0).astype(np.float64) + p.max(0)) / 2.0 for p in sps])
    got = S.chamfer.create_cd(sps, cents)
    want = oracle.create_cd(sps, cents)
    assert np.array_equal(got, want)
    assert np.array_equal(S.chamfer.farthest_superpoint_sample(sps, cents, 4, 1),
                          oracle.farthest_superpoint_sample(sps, cents, 4, 1))
