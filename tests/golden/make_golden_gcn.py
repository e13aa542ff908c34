"""Generate tests/golden/gcn.npz by running the UNMODIFIED reference fps_gcn_cpu.fps_adj_all / GCN_FPS_sampling in this
container on a small synthetic data set written with the reference's own helper_ply.write_ply.

Run:  python tests/golden/make_golden_gcn.py      (needs /root/reference; the fixture is committed)

The reference spells the dtype `np.float` (fps_gcn_cpu.py:64-65), an alias numpy removed in 1.24: it is restored for
the duration of the run (an environment shim, the reference source is not touched).
"""
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/SSDR_AL_s3dis")
if not hasattr(np, "float"):
    np.float = float  # noqa: NPY001
import fps_gcn_cpu  # noqa: E402
from helper_ply import write_ply  # noqa: E402


def make_room(rng, n_sp):
    """A room of n_sp superpoints: blobs of 40..160 points each, float32 coordinates like a .ply holds."""
    xyz, comps = [], []
    start = 0
    for _ in range(n_sp):
        n = int(rng.integers(40, 160))
        centre = rng.random(3) * np.array([6.0, 4.0, 2.5])
        pts = centre + rng.normal(0, 0.15, (n, 3)) * np.array([1.0, 1.0, 0.3])
        xyz.append(pts)
        comps.append(list(range(start, start + n)))
        start += n
    return np.concatenate(xyz).astype(np.float32), comps


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "superpoint"))
        os.makedirs(os.path.join(tmp, "input"))
        rooms = {"Area_1_office_1": 9, "Area_1_hallway_2": 7, "Area_2_storage_1": 5}
        refs = []
        for name, n_sp in rooms.items():
            xyz, comps = make_room(rng, n_sp)
            write_ply(os.path.join(tmp, "input", name + ".ply"), [xyz], ["x", "y", "z"])
            with open(os.path.join(tmp, "data", "superpoint", name + ".superpoint"), "wb") as f:
                pickle.dump({"components": comps}, f)
            out[name + "_xyz"] = xyz
            out[name + "_sizes"] = np.array([len(c) for c in comps], np.int64)
            refs += [{"cloud_name": name, "sp_idx": i} for i in range(n_sp)]
        order = rng.permutation(len(refs))
        unl = [refs[i] for i in order[:15]]
        lab = [refs[i] for i in order[15:]]
        out["room_names"] = np.array(list(rooms))
        out["unlabeled_cloud"] = np.array([r["cloud_name"] for r in unl])
        out["unlabeled_sp"] = np.array([r["sp_idx"] for r in unl], np.int64)
        out["labeled_cloud"] = np.array([r["cloud_name"] for r in lab])
        out["labeled_sp"] = np.array([r["sp_idx"] for r in lab], np.int64)
        fu = rng.standard_normal((len(unl), 13))
        fl = rng.standard_normal((len(lab), 13))
        out["unlabeled_features"], out["labeled_features"] = fu, fl
        adj, _ = fps_gcn_cpu.fps_adj_all(lab, unl, os.path.join(tmp, "input"), os.path.join(tmp, "data"))
        out["adj"] = adj
        for g, top in ((1, 0), (2, 0), (1, 4)):
            # the part of GCN_FPS_sampling between the adjacency and the FPS call (fps_gcn_cpu.py:153-167), run by
            # executing the reference function with farthest_features_sample replaced by a recorder
            seen = {}

            def recorder(features, n):
                seen["features"] = np.array(features)
                return np.arange(n)

            keep = fps_gcn_cpu.farthest_features_sample
            fps_gcn_cpu.farthest_features_sample = recorder
            try:
                fps_gcn_cpu.GCN_FPS_sampling(fl, lab, fu, unl, os.path.join(tmp, "input"), os.path.join(tmp, "data"), 5, g, top)
            finally:
                fps_gcn_cpu.farthest_features_sample = keep
            out["combo_g%d_top%d" % (g, top)] = seen["features"]
        # the whole function with its own FPS, first pick pinned through numpy's global seed
        np.random.seed(7)
        fl_sel = fps_gcn_cpu.GCN_FPS_sampling(fl, lab, fu, unl, os.path.join(tmp, "input"), os.path.join(tmp, "data"), 6, 1, 0)
        out["selected_clouds"] = np.array([k for k, v in fl_sel.items() for _ in v])
        out["selected_sp"] = np.array([s for k, v in fl_sel.items() for s in v], np.int64)
    np.savez_compressed(os.path.join(HERE, "gcn.npz"), **out)
    print("wrote gcn.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
