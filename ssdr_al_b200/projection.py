"""1-NN projection of the raw cloud onto its sub-sampled cloud: the step that follows grid
subsampling in every data-prep script, e.g. utils/data_prepare_s3dis.py:66-72

    search_tree = KDTree(sub_xyz); proj_idx = np.squeeze(search_tree.query(xyz, return_distance=False)); proj_idx.astype(np.int32)

The reference uses sklearn's KDTree (float64 arithmetic); here it is the same K=1 kernel as the up-sampling queries,
in float32 with the nanoflann operation order.  Indices agree with sklearn wherever the nearest sub-point is unique in
both precisions (checked in tests/test_projection_gpu.py)."""
import numpy as np

from . import nearest_neighbors


def project(xyz, sub_xyz):
    """(N,3) raw points, (M,3) sub-sampled points -> (N,) int32 index of the nearest sub-sampled point."""
    idx = nearest_neighbors.knn(np.asarray(sub_xyz), np.asarray(xyz), 1, omp=True)
    return np.squeeze(idx, axis=1).astype(np.int32)
