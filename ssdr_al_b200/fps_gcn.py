"""The superpoint adjacency and the feature propagation in front of the FPS loop, behind the reference's interface
(fps_gcn_cpu.py:40-117 `fps_adj_all`, :150-178 `GCN_FPS_sampling`).

The reference assembles two dense N x N float64 matrices on the host (centre distances and chamfer distances of the
superpoints that share a room, 1e10 elsewhere), turns them into a normalised adjacency `D^-1 (exp(-(ed+cd)) - I) + I`,
multiplies the superpoint features by it `gcn_number` times and hands the sum of the products to
`farthest_features_sample`.  Here the chamfer blocks come from `chamfer.create_cd` (one CUDA kernel per room), the
matrix is assembled, normalised and multiplied on the device (`csrc/gcn.cu`), and it never leaves the device unless
the caller asks for it.
"""
import ctypes as C
import pickle
import time
from os.path import join

import numpy as np

from . import _lib
from .chamfer import create_cd
from .selection import farthest_features_sample


class Adjacency(object):
    """The normalised adjacency of `fps_adj_all`, resident on the device.  `numpy()` copies it to the host."""

    def __init__(self, handle, n):
        self.handle = handle
        self.shape = (n, n)

    def numpy(self):
        out = np.empty(self.shape, np.float64)
        _lib.check(_lib.lib().ssdr_gcn_fetch(self.handle, _lib.ptr(out)))
        return out

    def close(self):
        if self.handle is not None:
            _lib.lib().ssdr_gcn_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


def adjacency_from_rooms(n_total, rooms):
    """rooms: iterable of (ref_idx_list, centre_xyz (n, 3), cd (n, n)) -- the per-room quantities of
    fps_gcn_cpu.py:77-101.  Returns the device-resident normalised adjacency (fps_gcn_cpu.py:104-116)."""
    refs, cents, cds, off = [], [], [], [0]
    for ref, centre, cd in rooms:
        ref = np.asarray(ref, np.int64).reshape(-1)
        n = len(ref)
        centre = np.asarray(centre, np.float64).reshape(n, 3)
        cd = np.asarray(cd, np.float64).reshape(n, n)
        refs.append(ref)
        cents.append(centre)
        cds.append(cd.reshape(-1))
        off.append(off[-1] + n)
    nb = len(refs)
    block_off = np.asarray(off, np.int64)
    ref = np.ascontiguousarray(np.concatenate(refs)) if nb else np.zeros(0, np.int64)
    cent = np.ascontiguousarray(np.concatenate(cents)) if nb else np.zeros((0, 3))
    cd = np.ascontiguousarray(np.concatenate(cds)) if nb else np.zeros(0)
    h = C.c_void_p()
    _lib.check(_lib.lib().ssdr_gcn_adjacency_f64(int(n_total), nb, _lib.ptr(block_off), _lib.ptr(ref), _lib.ptr(cent),
                                                 _lib.ptr(cd), C.byref(h)))
    return Adjacency(h, int(n_total))


def _group_by_cloud(labeled_select_ref, unlabeled_candidate_ref):
    """fps_gcn_cpu.py:43-61: rows 0 .. U-1 are the unlabeled candidates, U .. U+L-1 the labeled superpoints; rooms in
    order of first appearance."""
    total_cloud, names = {}, []
    for i, r in enumerate(unlabeled_candidate_ref):
        if r["cloud_name"] not in total_cloud:
            total_cloud[r["cloud_name"]] = []
            names.append(r["cloud_name"])
        total_cloud[r["cloud_name"]].append({"sp_idx": r["sp_idx"], "ref_idx": i})
    u = len(unlabeled_candidate_ref)
    for i, r in enumerate(labeled_select_ref):
        if r["cloud_name"] not in total_cloud:
            total_cloud[r["cloud_name"]] = []
            names.append(r["cloud_name"])
        total_cloud[r["cloud_name"]].append({"sp_idx": r["sp_idx"], "ref_idx": u + i})
    return total_cloud, names


def fps_adj_device(labeled_select_ref, unlabeled_candidate_ref, input_path, data_path, read_ply=None):
    """fps_adj_all with the result left on the device (what GCN_FPS_sampling needs)."""
    if read_ply is None:
        from helper_ply import read_ply  # the reference's own reader (fps_gcn_cpu.py:8), on the caller's path
    total_cloud, names = _group_by_cloud(labeled_select_ref, unlabeled_candidate_ref)
    n_total = len(unlabeled_candidate_ref) + len(labeled_select_ref)
    rooms = []
    for cloud_name in names:
        with open(join(data_path, "superpoint", cloud_name + ".superpoint"), "rb") as f:
            components = pickle.load(f)["components"]
        data = read_ply(join(input_path, "{:s}.ply".format(cloud_name)))
        xyz = np.vstack((data["x"], data["y"], data["z"])).T
        members = total_cloud[cloud_name]
        centre = np.zeros([len(members), 3])
        superpoints, ref = [], []
        for j, m in enumerate(members):
            x_y_z = xyz[components[m["sp_idx"]]]
            ref.append(m["ref_idx"])
            # the centre of the bounding box, in the coordinates' own dtype first (fps_gcn_cpu.py:86-88)
            centre[j, 0] = (np.min(x_y_z[:, 0]) + np.max(x_y_z[:, 0])) / 2.0
            centre[j, 1] = (np.min(x_y_z[:, 1]) + np.max(x_y_z[:, 1])) / 2.0
            centre[j, 2] = (np.min(x_y_z[:, 2]) + np.max(x_y_z[:, 2])) / 2.0
            superpoints.append(x_y_z)
        rooms.append((ref, centre, create_cd(superpoints, centre)))
    return adjacency_from_rooms(n_total, rooms)


def fps_adj_all(labeled_select_ref, unlabeled_candidate_ref, input_path, data_path, read_ply=None):
    """fps_gcn_cpu.py:40-117: returns (adj (N, N) float64 numpy array, seconds)."""
    begin_time = time.time()
    a = fps_adj_device(labeled_select_ref, unlabeled_candidate_ref, input_path, data_path, read_ply=read_ply)
    adj = a.numpy()
    a.close()
    return adj, time.time() - begin_time


def propagate(adj, features_v, gcn_number, gcn_top=0):
    """fps_gcn_cpu.py:153-167: the sum of V, adj V, adj (adj V), ... (gcn_number products) with the optional top-k mask
    of every row of adj.  adj: an `Adjacency` (device) or a (N, N) float64 numpy array."""
    v = np.ascontiguousarray(features_v, dtype=np.float64)
    n, d = v.shape
    out = np.empty((n, d), np.float64)
    top = int(gcn_top) if gcn_top > 0 else 0
    if isinstance(adj, Adjacency):
        _lib.check(_lib.lib().ssdr_gcn_propagate_f64(adj.handle, None, n, _lib.ptr(v), d, int(gcn_number), top,
                                                     _lib.ptr(out)))
    else:
        a = np.ascontiguousarray(adj, dtype=np.float64)
        if a.shape != (n, n):
            raise ValueError("adj must be (%d, %d), got %r" % (n, n, a.shape))
        _lib.check(_lib.lib().ssdr_gcn_propagate_f64(None, _lib.ptr(a), n, _lib.ptr(v), d, int(gcn_number), top,
                                                     _lib.ptr(out)))
    return out


def GCN_FPS_sampling(labeled_select_features, labeled_select_ref, unlabeled_candidate_features, unlabeled_candidate_ref,
                     input_path, data_path, sampling_batch, gcn_number, gcn_top, read_ply=None):
    """fps_gcn_cpu.py:150-178, same arguments and the same {cloud_name: [sp_idx, ...]} result."""
    adj = fps_adj_device(labeled_select_ref, unlabeled_candidate_ref, input_path, data_path, read_ply=read_ply)
    features_v = np.concatenate([unlabeled_candidate_features, labeled_select_features])
    combinational_features = propagate(adj, features_v, int(gcn_number), gcn_top)
    adj.close()
    unlabeled_num = len(unlabeled_candidate_features)
    selected_ids = farthest_features_sample(combinational_features[:unlabeled_num], sampling_batch)
    file_list = {}
    for i in selected_ids:
        cloud_name, sp_idx = unlabeled_candidate_ref[i]["cloud_name"], unlabeled_candidate_ref[i]["sp_idx"]
        if cloud_name not in file_list:
            file_list[cloud_name] = []
        file_list[cloud_name].append(sp_idx)
    return file_list
