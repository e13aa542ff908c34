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
