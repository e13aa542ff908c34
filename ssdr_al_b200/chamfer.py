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
