"""GPU: raw -> sub-sampled 1-NN projection (the step after grid subsampling in the data-prep scripts) against
sklearn's KDTree, which is what the reference calls (utils/data_prepare_s3dis.py:66-72)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_projection_matches_sklearn_kdtree():
    from sklearn.neighbors import KDTree
    import ssdr_al_b200 as S
    from ssdr_al_b200 import projection
    rng = np.random.default_rng(4)
    n = 400_000
    xyz = (rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])).astype(np.float32)
    xyz[: n // 2, 2] = rng.normal(0, 0.005, n // 2)
    sub = S.grid_subsampling.compute(xyz, sampleDl=0.04)
    got = projection.project(xyz, sub)
    assert got.dtype == np.int32 and got.shape == (n,)
    tree = KDTree(sub)
    dist, want = tree.query(xyz, k=2)
    want = want[:, 0].astype(np.int32)
    differ = np.flatnonzero(got != want)
    # a disagreement is only legitimate when float32 cannot separate the two nearest candidates
    d_got = np.sqrt(((xyz[differ].astype(np.float64) - sub[got[differ]].astype(np.float64)) ** 2).sum(1))
    assert np.all(d_got <= dist[differ, 0] * (1 + 1e-5) + 1e-7)
    assert differ.size < 1e-4 * n
