"""GPU parity: grid subsampling through the C ABI vs golden vectors (made by the reference) and the oracle.
Voxel membership, counts, barycentres, mean features and label votes are compared BIT-EXACTLY; rows are aligned by
voxel key because the product emits ascending-key order while the reference emits libstdc++ hash order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ssdr_al_b200 as S
    return S.grid_subsampling


def _room(rng, n):
    face = rng.integers(0, 3, n)
    p = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
    p[face == 0, 2] = 0.0
    p[face == 1, 1] = 0.0
    p[face == 2, 0] = 0.0
    p += rng.normal(0, 0.005, p.shape)
    return p.astype(np.float32)


def _by_key(oracle, pts, dl, ref_rows):
    """Re-order reference-order rows (tuple of arrays) into ascending voxel-key order."""
    o = oracle.grid_subsample(pts, None, None, dl, order="reference", with_keys=True)
    perm = np.argsort(o[3], kind="stable")
    assert o[0].tobytes() == ref_rows[0].tobytes()  # the fixture rows are in reference order
    return [None if r is None else r[perm] for r in ref_rows], o[3][perm], o[4][perm]


def test_grid_golden_full(G, oracle, golden):
    g = golden.grid
    (wp, wf, wc), wkeys, wcounts = _by_key(oracle, g["pts"], 0.1, (g["out_pts"], g["out_rgb"], g["out_lab"]))
    (p, f, c), keys, counts = G.compute(g["pts"], features=g["rgb"], classes=g["lab"], sampleDl=0.1, return_keys=True)
    assert p.dtype == np.float32 and f.dtype == np.float32 and c.dtype == np.int32 and c.shape == (len(p), 1)
    assert np.array_equal(keys, wkeys) and np.array_equal(counts, wcounts)
    assert p.tobytes() == wp.tobytes()
    assert f.tobytes() == wf.tobytes()
    assert np.array_equal(c, wc)


def test_grid_golden_points_only_negative(G, oracle, golden):
    g = golden.grid
    (wp,), _, _ = _by_key(oracle, g["pts"] - 3.0, 0.04, (g["neg_out_pts"],))
    p = G.compute(g["pts"] - 3.0, sampleDl=0.04)
    assert isinstance(p, np.ndarray) and p.tobytes() == wp.tobytes()


def test_grid_golden_label_collisions_two_columns(G, oracle, golden):
    g = golden.grid
    (wp, wc), _, _ = _by_key(oracle, g["pts"], 0.25, (g["lab2_out_pts"], g["lab2_out_lab"]))
    p, c = G.compute(g["pts"], classes=g["lab2"], sampleDl=0.25)
    assert p.tobytes() == wp.tobytes()
    assert c.shape == wc.shape and np.array_equal(c, wc)


@pytest.mark.parametrize("n,dl,fdim", [(200_000, 0.04, 3), (50_000, 0.5, 5), (30_000, 0.02, 1), (1000, 10.0, 3)])
def test_grid_vs_oracle(G, oracle, n, dl, fdim):
    rng = np.random.default_rng(n)
    pts = _room(rng, n)
    feats = rng.integers(0, 256, (n, fdim)).astype(np.uint8)
    lab = rng.integers(0, 13, n).astype(np.uint8)
    wp, wf, wc, wk, wn = oracle.grid_subsample(pts, feats, lab, dl, order="key", with_keys=True)
    (p, f, c), k, cnt = G.compute(pts, features=feats, classes=lab, sampleDl=dl, return_keys=True)
    assert np.array_equal(k, wk) and np.array_equal(cnt, wn)
    assert p.tobytes() == wp.tobytes() and f.tobytes() == wf.tobytes() and np.array_equal(c, wc)


@pytest.mark.parametrize("fdt,fdim,ldt,ldim", [("f4", 3, "i4", 1), ("u1", 3, "u1", 1), ("u1", 3, "i4", 1), ("f4", 5, "u1", 2),
                                               ("u1", 6, None, 0), (None, 0, "i4", 1), ("f4", 1, "i4", 3)])
def test_grid_packed_records_equal_plain_gather(G, oracle, monkeypatch, fdt, fdim, ldt, ldim):
    """Clouds beyond the L2 are reduced from packed (xyz | features | labels) records; force that path on a small cloud
    for every input layout (record sizes 16 .. 48 bytes, labels at aligned and unaligned offsets, a voxel heavy enough
    for the one-CTA-per-voxel reduce) and compare with the oracle and with the plain gather bit for bit."""
    rng = np.random.default_rng(fdim * 10 + ldim)
    n = 60_000
    pts = _room(rng, n)
    pts[:3000] = pts[0] + (rng.random((3000, 3)) * 0.01).astype(np.float32)  # > 1024 points in one voxel
    feats = None if fdt is None else (rng.integers(0, 256, (n, fdim)).astype(np.uint8) if fdt == "u1"
                                      else rng.random((n, fdim), dtype=np.float32))
    lab = None
    if ldt is not None:  # piecewise-constant labels with 3 % noise: single-label voxels (one thread each) and votes
        lab = np.repeat((3 * (pts[:, 0] > 3.5) + 5 * (pts[:, 1] > 2.0))[:, None], ldim, 1)
        noisy = rng.random((n, ldim)) < 0.03
        lab = np.where(noisy, rng.integers(0, 13, (n, ldim)), lab).astype(np.uint8 if ldt == "u1" else np.int32)
    want = oracle.grid_subsample(pts, feats, lab, 0.05, order="key", with_keys=True)
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SSDR_GRID_PACK", mode)
        got[mode] = G.compute(pts, features=feats, classes=lab, sampleDl=0.05, return_keys=True)
    monkeypatch.setenv("SSDR_GRID_PACK", "1")
    monkeypatch.setenv("SSDR_GRID_SMALL", "0")  # every voxel by the eight-lane groups (no thread-per-voxel pass)
    got["groups"] = G.compute(pts, features=feats, classes=lab, sampleDl=0.05, return_keys=True)
    for mode in ("1", "0", "groups"):
        res, k, cnt = got[mode]
        res = res if isinstance(res, tuple) else (res,)
        assert np.array_equal(k, want[3]) and np.array_equal(cnt, want[4])
        assert res[0].tobytes() == want[0].tobytes()
        j = 1
        if feats is not None:
            assert res[j].tobytes() == want[1].tobytes()
            j += 1
        if lab is not None:
            assert np.array_equal(res[j], want[2])


def test_grid_heavy_voxel_and_many_labels(G, oracle):
    rng = np.random.default_rng(11)
    n = 40_000
    pts = (rng.random((n, 3)) * 0.05).astype(np.float32)  # everything in a handful of voxels
    pts[:100] += 5.0
    lab = rng.integers(-20, 30, n).astype(np.int32)  # 50 distinct labels -> inner map rehashes 13 -> 29 -> 59
    wp, _, wc = oracle.grid_subsample(pts, None, lab, 0.04, order="key")
    p, c = G.compute(pts, classes=lab, sampleDl=0.04)
    assert p.tobytes() == wp.tobytes() and np.array_equal(c, wc)


def test_grid_wide_extent_needs_more_than_32_key_bits(G, oracle):
    rng = np.random.default_rng(12)
    n = 100_000
    pts = (rng.random((n, 3)) * np.array([200.0, 200.0, 30.0])).astype(np.float32)
    pts[: n // 2] = (rng.random((n // 2, 3)) * 2.0).astype(np.float32)  # dense near the "scanner"
    wp, _, _, wk, wn = oracle.grid_subsample(pts, None, None, 0.06, order="key", with_keys=True)
    assert int(wk.max()) > 2 ** 32
    p, k, cnt = G.compute(pts, sampleDl=0.06, return_keys=True)
    assert np.array_equal(k, wk) and np.array_equal(cnt, wn) and p.tobytes() == wp.tobytes()


def test_grid_wrapper_contract(G):
    pts = np.random.default_rng(1).random((100, 3))
    assert isinstance(G.compute(pts, sampleDl=0.2), np.ndarray)
    r = G.compute(pts, features=np.ones((100, 2)), sampleDl=0.2)
    assert isinstance(r, tuple) and len(r) == 2 and r[1].shape[1] == 2
    r = G.compute(pts, classes=np.zeros((100, 2), np.int64), sampleDl=0.2)
    assert len(r) == 2 and r[1].shape[1] == 2 and r[1].dtype == np.int32
    with pytest.raises(RuntimeError, match="points.shape is not"):
        G.compute(np.zeros((10, 2)), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="features.shape is not"):
        G.compute(pts, features=np.ones((99, 2)), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="classes.shape is not"):
        G.compute(pts, classes=np.zeros(7, np.int32), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="Error parsing method"):
        G.compute(pts, sampleDl=0.1, method="nope")
    with pytest.raises(TypeError):
        G.compute(pts, np.ones((100, 3)))  # everything after `points` is keyword-only (format "O|$OOfsi")
    with pytest.raises(RuntimeError, match="^Error$"):
        G.compute(np.zeros((0, 3)), sampleDl=0.1)


def test_grid_config1_full_size_properties(G):
    """BASELINE config 1 size (1M pts): counts sum to N, keys strictly ascending, features/labels in range."""
    rng = np.random.default_rng(0)
    n = 1_000_000
    pts = _room(rng, n)
    pts -= pts.min(0)
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    lab = (pts[:, 0] * 1.7).astype(np.uint8) % 13
    (p, f, c), k, cnt = G.compute(pts, features=rgb, classes=lab, sampleDl=0.04, return_keys=True)
    assert cnt.sum() == n and (np.diff(k.astype(np.int64)) > 0).all()
    assert f.min() >= 0 and f.max() <= 255 and c.min() >= 0 and c.max() < 13
    # subsampling the barycentres at the same cell size can only merge, never split
    p2, k2, cnt2 = G.compute(p, sampleDl=0.04, return_keys=True)
    assert len(p2) <= len(p) and cnt2.sum() == len(p)


def test_grid_reference_row_order_golden(G, golden):
    """order="reference": rows come out in the reference's own libstdc++ hash-iteration order -- the fixture rows,
    produced by the unmodified reference, must be reproduced bit for bit INCLUDING their order."""
    g = golden.grid
    p, f, c = G.compute(g["pts"], features=g["rgb"], classes=g["lab"], sampleDl=0.1, order="reference")
    assert p.tobytes() == g["out_pts"].tobytes()
    assert f.tobytes() == g["out_rgb"].tobytes()
    assert np.array_equal(c, g["out_lab"])
    p = G.compute(g["pts"] - 3.0, sampleDl=0.04, order="reference")
    assert p.tobytes() == g["neg_out_pts"].tobytes()
    p, c = G.compute(g["pts"], classes=g["lab2"], sampleDl=0.25, order="reference")
    assert p.tobytes() == g["lab2_out_pts"].tobytes() and np.array_equal(c, g["lab2_out_lab"])


@pytest.mark.parametrize("n,dl", [(300_000, 0.04), (50_000, 0.3), (20, 0.5), (2_000_000, 0.05)])
def test_grid_reference_row_order_vs_oracle(G, oracle, n, dl):
    rng = np.random.default_rng(n + 1)
    pts = _room(rng, n)
    lab = rng.integers(0, 13, n).astype(np.uint8)
    wp, _, wc = oracle.grid_subsample(pts, None, lab, dl, order="reference")
    p, c = G.compute(pts, classes=lab, sampleDl=dl, order="reference")
    assert p.shape == wp.shape and p.tobytes() == wp.tobytes() and np.array_equal(c, wc)


# ---- slab ownership (multi-GPU sharding of one cloud), exercised with virtual ranks on one GPU ----------------
def _slab_cloud(rng, n):
    p = (rng.random((n, 3)) * np.array([40.0, 30.0, 6.0])).astype(np.float32)
    p[: n // 3, 2] = rng.normal(1.0, 0.01, n // 3).astype(np.float32)  # a dense floor: unbalanced layers
    p[-2500:] = p[-1] + (rng.random((2500, 3)) * 0.02).astype(np.float32)  # one voxel heavy enough for the one-CTA reduce
    p += np.float32(-7.3)
    f = rng.random((n, 3)).astype(np.float32)
    c = rng.integers(0, 13, (n, 1)).astype(np.int32)
    return p, f, c


@pytest.mark.parametrize("world", [2, 3, 5])
def test_grid_slabs_replicated_cloud_equal_single_run(G, world, monkeypatch):
    """Every virtual rank sees the whole cloud and reduces only its layers: rows concatenated in rank order are the
    single-run result, bit for bit (points, features, labels, keys, counts)."""
    import torch
    from ssdr_al_b200 import device as dev, dist as SD
    rng = np.random.default_rng(40 + world)
    p, f, c = _slab_cloud(rng, 300_000)
    tp, tf, tc = (torch.from_numpy(a).cuda() for a in (p, f, c))
    want = dev.grid_subsample(tp, tf, tc, 0.11, return_keys=True)
    layers, n_layers = dev.grid_point_layers(tp, 0.11, 2)
    assert n_layers == int(layers.max()) + 1
    bounds = SD.balanced_slabs(torch.bincount(layers.long(), minlength=n_layers).cpu().numpy(), world)
    parts = [dev.grid_subsample(tp, tf, tc, 0.11, slab=(2, int(bounds[r]), int(bounds[r + 1])), return_keys=True)
             for r in range(world)]
    sizes = [len(x[3]) for x in parts]
    assert sum(sizes) == len(want[3]) and max(sizes) < len(want[3])  # a rank may own nothing (one dense layer)
    for j in range(3):
        assert torch.equal(torch.cat([x[j] for x in parts]), want[j])
    assert np.array_equal(np.concatenate([x[3] for x in parts]), want[3])
    assert np.array_equal(np.concatenate([x[4] for x in parts]), want[4])
    # the slab members are reduced from packed records by default; the gather through the input index gives the same
    monkeypatch.setenv("SSDR_GRID_PACK", "0")
    for r in range(world):
        x = dev.grid_subsample(tp, tf, tc, 0.11, slab=(2, int(bounds[r]), int(bounds[r + 1])), return_keys=True)
        assert all(torch.equal(x[j], parts[r][j]) for j in range(3)) and np.array_equal(x[3], parts[r][3])
    monkeypatch.delenv("SSDR_GRID_PACK")
    # an empty slab is a valid shard
    e = dev.grid_subsample(tp, tf, tc, 0.11, slab=(2, n_layers + 5, n_layers + 9), return_keys=True)
    assert e[0].shape == (0, 3) and e[1].shape == (0, 3) and len(e[3]) == 0
    # slabs along x partition the voxels as well (only the concatenation order differs)
    lx, nx = dev.grid_point_layers(tp, 0.11, 0)
    bx = SD.balanced_slabs(torch.bincount(lx.long(), minlength=nx).cpu().numpy(), world)
    px = [dev.grid_subsample(tp, tf, tc, 0.11, slab=(0, int(bx[r]), int(bx[r + 1])), return_keys=True)
          for r in range(world)]
    k = np.concatenate([x[3] for x in px])
    o = np.argsort(k, kind="stable")
    assert np.array_equal(k[o], want[3])
    assert torch.equal(torch.cat([x[0] for x in px])[torch.from_numpy(o).cuda()], want[0])


@pytest.mark.parametrize("world", [2, 4])
def test_grid_slabs_row_chunks_routed_to_owners_equal_single_run(G, world):
    """Row-sharded input: chunk bboxes combined, points routed to their slab owner in input order (what the
    all-to-all of dist.slab_exchange delivers), each owner run with the whole cloud's bbox."""
    import torch
    from ssdr_al_b200 import device as dev, dist as SD
    rng = np.random.default_rng(50 + world)
    p, f, c = _slab_cloud(rng, 200_000)
    tp, tf, tc = (torch.from_numpy(a).cuda() for a in (p, f, c))
    want = dev.grid_subsample(tp, tf, tc, 0.09, return_keys=True)
    spans = [SD.shard_range(len(p), world, r) for r in range(world)]
    boxes = np.array([dev.grid_bbox(tp[b:e]) for b, e in spans], np.float32)
    bbox = np.concatenate([boxes[:, :3].min(0), boxes[:, 3:].max(0)])
    assert np.array_equal(bbox, np.concatenate([p.min(0), p.max(0)]))
    lay = [dev.grid_point_layers(tp[b:e].contiguous(), 0.09, 2, bbox) for b, e in spans]
    n_layers = max(n for _, n in lay)
    assert n_layers == dev.grid_layers(bbox, 0.09, 2)
    layers = torch.cat([l for l, _ in lay]).long()
    # layer histogram and the stable partition by slab owner, both done by the library (what dist.slab_exchange uses)
    hist = sum(dev.grid_layer_hist(tp[b:e].contiguous(), 0.09, 2, bbox) for b, e in spans)
    assert torch.equal(hist, torch.bincount(layers, minlength=n_layers))
    bounds = SD.balanced_slabs(hist.cpu().numpy(), world)
    routed = [dev.grid_route(tp[b:e], tf[b:e], tc[b:e], 0.09, 2, bbox, bounds) for b, e in spans]
    dest = torch.searchsorted(torch.from_numpy(bounds[1:-1].copy()).cuda(), layers, right=True)
    parts = []
    for r in range(world):
        # what the all-to-all delivers to rank r: group r of every source chunk, in source-rank order
        pieces = [[], [], []]
        for (rp, rf, rc, cnt) in routed:
            off = sum(cnt[:r])
            for j, t in enumerate((rp, rf, rc)):
                pieces[j].append(t[off:off + cnt[r]])
        got = [torch.cat(x) for x in pieces]
        m = dest == r
        assert torch.equal(got[0], tp[m]) and torch.equal(got[1], tf[m]) and torch.equal(got[2], tc[m])
        parts.append(dev.grid_subsample(got[0], got[1], got[2], 0.09, bbox=bbox, return_keys=True))
    for j in range(3):
        assert torch.equal(torch.cat([x[j] for x in parts]), want[j])
    assert np.array_equal(np.concatenate([x[3] for x in parts]), want[3])
    assert np.array_equal(np.concatenate([x[4] for x in parts]), want[4])


def test_grid_slab_argument_errors(G):
    import ctypes as C
    import torch
    from ssdr_al_b200 import _lib, device as dev
    tp = torch.rand((100, 3), device="cuda")
    with pytest.raises(RuntimeError, match="bbox min exceeds max"):
        dev.grid_subsample(tp, sampleDl=0.1, bbox=[1, 0, 0, 0, 1, 1])
    M, h = C.c_size_t(0), C.c_void_p()
    rc = _lib.lib().ssdr_grid_subsample_slab_dev(C.c_void_p(tp.data_ptr()), None, None, 100, 0, 0, 0.1, 1, None, 2,
                                                 0, 5, None, C.byref(M), C.byref(h))
    assert rc != 0 and b"whole cloud only" in _lib.lib().ssdr_last_error()


def test_grid_uint8_inputs_widened_on_device_equal_host_conversion(G):
    """uint8 colours / labels (what the reference's callers pass) are uploaded as bytes and widened on the device;
    the result must be the bits of the host-converted float32 / int32 call, for 1 and 2 label columns."""
    rng = np.random.default_rng(12)
    p = _room(rng, 120_000)
    rgb = rng.integers(0, 256, (len(p), 3)).astype(np.uint8)
    for lab in (rng.integers(0, 13, len(p)).astype(np.uint8), rng.integers(0, 200, (len(p), 2)).astype(np.uint8)):
        a = G.compute(p, features=rgb, classes=lab, sampleDl=0.05)
        b = G.compute(p, features=rgb.astype(np.float32), classes=lab.astype(np.int32), sampleDl=0.05)
        assert all(x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes() for x, y in zip(a, b))
        c = G.compute(p, features=rgb, classes=lab.astype(np.int32), sampleDl=0.05)  # mixed
        assert all(x.tobytes() == y.tobytes() for x, y in zip(a, c))
    only = G.compute(p, classes=rng.integers(0, 13, len(p)).astype(np.uint8), sampleDl=0.05)
    assert only[1].dtype == np.int32 and only[1].shape[1] == 1
