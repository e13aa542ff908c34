"""CPU: live comparison of the oracle restatement with the UNMODIFIED reference compiled as oracle/_ref/*.so
(built by oracle/Makefile where /root/reference exists; the prebuilt .so files travel to the GPU box)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def _clouds(rng, n):
    yield "uniform", rng.random((n, 3), dtype=np.float32)
    yield "quantised", (np.round(rng.random((n, 3)) * [300, 300, 2]) / 100).astype(np.float32)
    yield "duplicates", rng.random((n // 8, 3), dtype=np.float32)[rng.integers(0, n // 8, n)]
    yield "collinear", np.stack([rng.random(n), np.zeros(n), np.zeros(n)], 1).astype(np.float32)
    yield "identical", np.ones((n, 3), np.float32)


@pytest.mark.parametrize("K", [1, 16])
def test_knn_live(K):
    rng = np.random.default_rng(5)
    for name, p in _clouds(rng, 6000):
        assert np.array_equal(O.knn(p, p, K, threads=4), O.ref_knn(p, p, K, omp=True)), name


def test_knn_batch_live():
    rng = np.random.default_rng(6)
    P = rng.random((4, 3000, 3), dtype=np.float32)
    assert np.array_equal(O.knn_batch(P, P, 16, threads=4), O.ref_knn_batch(P, P, 16, omp=True))
    Q = np.ascontiguousarray(P[:, :750])
    assert np.array_equal(O.knn_batch(Q, P, 1, threads=4), O.ref_knn_batch(Q, P, 1, omp=False))


@pytest.mark.parametrize("dl", [0.04, 0.11])
def test_grid_live(dl):
    rng = np.random.default_rng(7)
    n = 60000
    p = (rng.random((n, 3)) * [7, 5, 3]).astype(np.float32)
    p[: n // 2, 2] = 0  # a dense floor
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    lab = rng.integers(0, 13, n).astype(np.uint8)
    a = O.grid_subsample(p, rgb, lab, dl)
    b = O.ref_grid_subsample(p, rgb, lab, dl)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()


@pytest.mark.skipif(not os.path.isdir("/root/reference/SSDR_AL_s3dis"), reason="reference checkout not present")
def test_chamfer_and_superpoint_fps_live():
    """Restatements vs the reference's own create_cd (fps_gcn_cpu.py) and farthest_superpoint_sample (sampler2.py, run
    from its source text) on fresh random superpoints: matrices within the KD tree's bound rounding, picks identical."""
    rng = np.random.default_rng(8)
    sps, cents = [], []
    for n in [int(v) for v in rng.integers(2, 260, 18)]:
        c = rng.random(3) * 4
        p = (c + rng.normal(0, 0.3, (n, 3)) * rng.choice([1.0, 0.03], 3)).astype(np.float32)
        sps.append(p)
        cents.append(np.asarray([(np.min(p[:, d]) + np.max(p[:, d])) / 2.0 for d in range(3)]))
    cents = np.asarray(cents)
    np.testing.assert_allclose(O.create_cd(sps, cents), O.ref_create_cd(sps, cents), rtol=1e-12, atol=0)
    assert np.array_equal(O.farthest_superpoint_sample(sps, cents, 9, 2),
                          O.ref_farthest_superpoint_sample(sps, cents, 9, 2))
    # float32 centroids, as the real caller passes them: the centroid term stays float32 in both
    c32 = [np.asarray(c, dtype=np.float32) for c in cents]
    assert np.array_equal(O.farthest_superpoint_sample(sps, c32, 9, 2),
                          O.ref_farthest_superpoint_sample(sps, c32, 9, 2))
