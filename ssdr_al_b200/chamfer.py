"""Chamfer adjacency behind the reference's `create_cd` interface (fps_gcn_cpu.py:25-38, fps_gcn_cuda.py): the
S x S matrix of symmetric mean nearest-neighbour distances between the centred superpoints of one room, which
`fps_adj_all` turns into the adjacency that feeds the FPS loop.  All pairs are evaluated by one CUDA kernel."""
import numpy as np

from . import _lib


def create_cd(superpoint_list, superpoint_centroid_list):
    """fps_gcn_cpu.py:25-38: centre every superpoint on its centroid (the same numpy subtraction, so the same float64
    values), then cd[c, i] = mean_i min_c |.| + mean_c min_i |.|, zeros on the diagonal.  Returns (S, S) float64."""
    sp_num = len(superpoint_list)
    if sp_num == 0:
        return np.zeros([0, 0])
    aligned = [np.asarray(superpoint_list[i] - superpoint_centroid_list[i], dtype=np.float64).reshape(-1, 3)
               for i in range(sp_num)]
    offsets = np.zeros(sp_num + 1, np.int64)
    np.cumsum([len(a) for a in aligned], out=offsets[1:])
    if (np.diff(offsets) == 0).any():
        raise RuntimeError("ssdr_al_b200.chamfer.create_cd: empty superpoint (the reference's KDTree rejects it too)")
    pts = np.ascontiguousarray(np.concatenate(aligned, axis=0))
    out = np.empty((sp_num, sp_num), np.float64)
    _lib.check(_lib.lib().ssdr_chamfer_matrix_f64(_lib.ptr(pts), _lib.ptr(offsets), sp_num, _lib.ptr(out)))
    return out


def farthest_superpoint_sample(superpoint_list, superpoint_centroid_list, sample_number, trigger_idx):
    """sampler2.py:49-80: farthest point sampling over superpoints, distance = squared centroid distance + chamfer
    distance to the current pick.  Returns (sample_number,) int32 starting with trigger_idx."""
    sp_num = len(superpoint_list)
    # the squared centroid distance is evaluated in the centroids' own dtype, like np.sum((cents - cur) ** 2, -1) does
    # (sampler2.py:68-69): float32 ply coordinates stay float32 until the float64 chamfer row is added
    cents = np.asarray(superpoint_centroid_list)
    cents = np.ascontiguousarray(cents if cents.dtype == np.float32 else cents.astype(np.float64)).reshape(sp_num, 3)
    aligned = [np.asarray(superpoint_list[i] - superpoint_centroid_list[i], dtype=np.float64).reshape(-1, 3)
               for i in range(sp_num)]
    offsets = np.zeros(sp_num + 1, np.int64)
    np.cumsum([len(a) for a in aligned], out=offsets[1:])
    if sp_num == 0 or (np.diff(offsets) == 0).any():
        raise RuntimeError("ssdr_al_b200.chamfer.farthest_superpoint_sample: no or empty superpoints")
    pts = np.ascontiguousarray(np.concatenate(aligned, axis=0))
    out = np.zeros(int(sample_number), np.int32)
    _lib.check(_lib.lib().ssdr_superpoint_fps(_lib.ptr(pts), _lib.ptr(offsets), sp_num, _lib.ptr(cents),
                                              0 if cents.dtype == np.float32 else 1, int(trigger_idx),
                                              int(sample_number), _lib.ptr(out)))
    return out
