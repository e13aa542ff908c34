"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

  KNN / grid subsampling : reference C++ compiled where it lies -> oracle/_ref/*.so (oracle/Makefile `ref`)
  FPS                    : /root/reference/SSDR_AL_s3dis/fps_gcn_cpu.py::farthest_features_sample (imported)
  k-center               : /root/reference/SSDR_AL_s3dis/kcenterGreedy.py::kCenterGreedy (imported; sklearn 1.9.0)

Run:  python tests/golden/make_golden.py      (needs /root/reference; the fixtures are committed)
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/SSDR_AL_s3dis")

from oracle import oracle as O  # noqa: E402

O.build(ref=True)
import fps_gcn_cpu  # noqa: E402
from kcenterGreedy import kCenterGreedy  # noqa: E402


def room(rng, n, quant=None):
    """Points on the floor / walls of a 7x5x3 m box (S3DIS-like surfaces)."""
    face = rng.integers(0, 3, n)
    p = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
    p[face == 0, 2] = 0.0
    p[face == 1, 1] = 0.0
    p[face == 2, 0] = 0.0
    p += rng.normal(0, 0.005, p.shape)
    if quant:
        p = np.round(p / quant) * quant
    return p.astype(np.float32)


def main():
    rng = np.random.default_rng(20261017)

    # ---- KNN ----
    knn = {}
    clouds = {
        "uniform": rng.random((3000, 3), dtype=np.float32),
        "room": room(rng, 3000),
        "quant": room(rng, 3000, quant=0.01),  # many exact fp32 distance ties
        "dups": rng.random((500, 3), dtype=np.float32)[rng.integers(0, 500, 3000)],  # data_aug-style duplicates
        "tiny": rng.random((9, 3), dtype=np.float32),  # single leaf
    }
    for name, p in clouds.items():
        knn[name + "_pts"] = p
        knn[name + "_k16"] = O.ref_knn(p, p, min(16, len(p))).astype(np.int32)
        q = (rng.random((700, 3)) * 1.3 - 0.15).astype(np.float32) * (p.max(0) - p.min(0)) + p.min(0)
        knn[name + "_q"] = q
        knn[name + "_q_k1"] = O.ref_knn(p, q, 1, omp=True).astype(np.int32)
        knn[name + "_q_k5"] = O.ref_knn(p, q, 5).astype(np.int32)
    bp = rng.random((3, 1200, 3), dtype=np.float32)
    bq = bp[:, :300, :]
    knn["batch_pts"] = bp
    knn["batch_k16"] = O.ref_knn_batch(bp, bp, 16, omp=True).astype(np.int32)
    knn["batch_sub_k1"] = O.ref_knn_batch(np.ascontiguousarray(bq), bp, 1).astype(np.int32)  # up-sampling 1-NN
    small = rng.random((7, 3), dtype=np.float32)
    knn["small_pts"] = small
    knn["small_k10"] = O.ref_knn(small, small[:1], 10).astype(np.int32)  # K > npts: first npts valid, rest 0
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **knn)

    # ---- grid subsampling ----
    grid = {}
    n = 20000
    p = room(rng, n)
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    lab = (p[:, 0] * 1.9).astype(np.uint8) % 13
    noisy = rng.random(n) < 0.3
    lab[noisy] = rng.integers(0, 13, int(noisy.sum()))
    grid["pts"], grid["rgb"], grid["lab"] = p, rgb, lab
    po, fo, co = O.ref_grid_subsample(p, rgb, lab, 0.1)
    grid["out_pts"], grid["out_rgb"], grid["out_lab"] = po, fo, co
    po, _, _ = O.ref_grid_subsample(p - 3.0, None, None, 0.04)  # negative coordinates, points only
    grid["neg_out_pts"] = po
    lab2 = np.stack([rng.integers(-5, 40, n), rng.integers(0, 3, n)], 1).astype(np.int32)  # collide mod 13, ldim 2
    grid["lab2"] = lab2
    po, _, co = O.ref_grid_subsample(p, None, lab2, 0.25)
    grid["lab2_out_pts"], grid["lab2_out_lab"] = po, co
    np.savez_compressed(os.path.join(HERE, "grid.npz"), **grid)

    # ---- FPS ----
    fps = {}
    for tag, shape, dt, picks in (("d32", (1500, 32), np.float32, 120), ("d256", (400, 256), np.float32, 60),
                                  ("d129_f64", (300, 129), np.float64, 60), ("d3", (2000, 3), np.float32, 100),
                                  ("d20", (600, 20), np.float32, 700)):
        F = rng.standard_normal(shape).astype(dt)
        np.random.seed(7)
        r = fps_gcn_cpu.farthest_features_sample(F, picks)
        fps[tag + "_F"] = F
        fps[tag + "_picks"] = r
    np.savez_compressed(os.path.join(HERE, "fps.npz"), **fps)

    # ---- k-center ----
    kc = {}
    for tag, shape, dt in (("d129_f64", (600, 129), np.float64), ("d32_f32", (1500, 32), np.float32),
                           ("d256_f64", (200, 256), np.float64)):
        X = rng.standard_normal(shape).astype(dt)
        sel = np.arange(shape[0] - 50, shape[0])
        with contextlib.redirect_stdout(io.StringIO()):
            r = kCenterGreedy(X).select_batch_(sel, 80)
        kc[tag + "_X"] = X
        kc[tag + "_sel"] = sel
        kc[tag + "_picks"] = np.asarray(r, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "kcenter.npz"), **kc)
    for f in ("knn", "grid", "fps", "kcenter"):
        print(f, os.path.getsize(os.path.join(HERE, f + ".npz")) // 1024, "KiB")


def superpoints(rng, sizes):
    """Superpoint-like clouds of one room: small float32 patches (planar, elongated, blobs, duplicates) and the bbox
    centres fps_adj_all computes for them (fps_gcn_cpu.py:84-88)."""
    sps, cents = [], []
    for n in sizes:
        c = rng.random(3) * np.array([7.0, 5.0, 3.0])
        p = c + rng.normal(0, 0.25, (n, 3)) * rng.choice([1.0, 0.02], 3)
        if n > 6:
            p[n // 2] = p[0]  # an exact duplicate
        p = p.astype(np.float32)
        sps.append(p)
        cents.append([(np.min(p[:, d]) + np.max(p[:, d])) / 2.0 for d in range(3)])
    return sps, np.array(cents)


def make_chamfer():
    """create_cd of the unmodified reference (fps_gcn_cpu.py:25-38, sklearn KDTree underneath)."""
    sys.path.insert(0, "/root/reference/SSDR_AL_s3dis")
    from fps_gcn_cpu import create_cd
    rng = np.random.default_rng(20)
    out = {}
    for tag, sizes in (("small", [1, 2, 7, 8, 9, 33, 127, 128, 129, 300]),
                       ("room", [int(v) for v in rng.integers(40, 900, 24)] + [2600])):
        sps, cents = superpoints(rng, sizes)
        out[tag + "_points"] = np.concatenate(sps)
        out[tag + "_offsets"] = np.concatenate([[0], np.cumsum([len(p) for p in sps])]).astype(np.int64)
        out[tag + "_centroids"] = cents
        out[tag + "_cd"] = create_cd(sps, cents)
        # farthest_superpoint_sample (sampler2.py:49-80), run from the reference's own source
        n_pick = min(len(sps), 10 if tag == "small" else 16)
        out[tag + "_fps_picks"] = O.ref_farthest_superpoint_sample(sps, cents, n_pick, 0 if tag == "small" else 3)
        out[tag + "_fps_trigger"] = np.int32(0 if tag == "small" else 3)
    np.savez_compressed(os.path.join(HERE, "chamfer.npz"), **out)
    print("chamfer", os.path.getsize(os.path.join(HERE, "chamfer.npz")) // 1024, "KiB")


if __name__ == "__main__":
    if sys.argv[1:] == ["chamfer"]:
        make_chamfer()  # added later with its own generator: the other fixtures stay byte-identical
    else:
        main()
        make_chamfer()
