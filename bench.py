#!/usr/bin/env python
"""bench.py -- BASELINE config 2: the RandLA-Net S3DIS input pyramid (knn_batch k=16 for 6 x 40,960 points plus the
four random-downsampled layers 10240/2560/640/160, and the 1-NN up-sampling query at every level), i.e. the loop of
SSDR_AL_s3dis/s3dis_dataset.py:164-177.

One "step" = one full pyramid for one batch of 6 synthetic clouds (seeded uniform box).
  value  = k=16 KNN queries/s with the clouds already resident in HBM (device-pointer C ABI, CUDA-event timed; the
           1-NN up-sampling work is inside the timed region but only k=16 queries are counted).
  e2e    = the same metric through the reference-facing Python API (`nearest_neighbors.knn_batch`, host numpy in,
           host int64 out), host<->device copies inside the timed region.
  roofline     = the dominant kernel (level-0 k=16 query kernel) against the measured HBM copy peak.
  cpu_baseline = the reference's own nanoflann/OpenMP code (oracle/_ref, or the C port when _ref is absent) timed on
                 this box's host cores on the same pyramid.
`--impl reference` runs only that CPU reference, as the driver's reference arm.
Multi-GPU (torchrun, one rank per GPU): batch items are independent, every rank runs its own batch of 6 clouds,
no data-path collective (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, N0, K = 6, 40960, 16
RATIOS = (4, 4, 4, 4, 2)
LEVELS = [N0 // d for d in (1, 4, 16, 64, 256)]  # 40960, 10240, 2560, 640, 160
K16_QUERIES = B * sum(LEVELS)                    # 327,360 per step
K1_QUERIES = B * sum(LEVELS)                     # up-sampling: queries = the level's points, support = next level
METRIC = "k16_knn_queries_per_s"
WORKLOAD = "randla_s3dis_pyramid_b6x40960_k16_plus_1nn_upsampling"


def make_clouds(seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, (B, N0, 3)) * np.array([2.0, 2.0, 1.5])).astype(np.float32)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    kernel (profiles/traffic.json, written by tools/ncu_summary.py traffic); None when no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return float(json.load(f)[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_pyramid(xyz, fn):
    """The loop of s3dis_dataset.py:164-177 with `fn(support, queries, k)` = knn_batch(omp=True)."""
    for ratio in RATIOS:
        fn(xyz, xyz, K)
        sub = np.ascontiguousarray(xyz[:, : xyz.shape[1] // ratio, :])
        fn(sub, xyz, 1)
        xyz = sub


def cpu_reference_fn():
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    if O.have_ref():
        return (lambda p, q, k: O.ref_knn_batch(p, q, k, omp=True)), "reference", min(B, cores)
    O.lib()
    t = min(B, cores)
    return (lambda p, q, k: O.knn_batch(p, q, k, threads=t)), "port", t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, kind, cores = cpu_reference_fn()
    clouds = make_clouds(1)
    for _ in range(args.warmup):
        cpu_pyramid(clouds, fn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pyramid(clouds, fn)
    dt = time.perf_counter() - t0
    val = K16_QUERIES * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch": B, "points": N0, "k": K, "levels": LEVELS},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": kind,
                         "sample": "%d full pyramid steps, knn_batch(omp=True): OpenMP over the %d batch items" %
                                   (args.steps, B)},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ssdr_al_b200 as S
    from ssdr_al_b200 import _lib, device as D

    dev = torch.device("cuda", local)
    host_clouds = make_clouds(1 + rank)
    xyz0 = torch.from_numpy(host_clouds).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # pre-allocated outputs so the timed region holds only our kernels and the slicing copies the caller does
    outs16 = [torch.zeros((B, n, K), dtype=torch.int64, device=dev) for n in LEVELS]
    outs1 = [torch.zeros((B, n, 1), dtype=torch.int64, device=dev) for n in LEVELS]

    def gpu_pyramid(collect=None):
        xyz = xyz0
        launches = 0
        for li, ratio in enumerate(RATIOS):
            r = D.knn_batch(xyz, xyz, K, out=outs16[li], want_stats=collect is not None)
            if collect is not None:
                collect.append(("k16", li, r[1]))
            sub = xyz[:, : xyz.shape[1] // ratio, :].contiguous()
            r1 = D.knn_batch(sub, xyz, 1, out=outs1[li], want_stats=collect is not None)
            if collect is not None:
                collect.append(("k1", li, r1[1]))
            xyz = sub
        return launches

    def count_launches(stats):
        # counted inside the library where the launches happen (memsets and torch's slicing copies not included)
        return sum(int(st["kernel_launches"]) for _, _, st in stats)

    for _ in range(max(args.warmup, 3)):
        gpu_pyramid()
    torch.cuda.synchronize()

    # one instrumented pass (untimed) for per-kernel timing, tie statistics and launch counts
    stats = []
    flush.zero_()
    gpu_pyramid(stats)
    torch.cuda.synchronize()
    launches_per_step = count_launches(stats)
    # dominant kernel, measured cold (L2 flushed) a few times
    dom_ms = []
    for _ in range(5):
        flush.zero_()
        _, st = D.knn_batch(xyz0, xyz0, K, out=outs16[0], want_stats=True)
        dom_ms.append(st["main_kernel_ms"])
    dom_ms_avg = float(np.mean(dom_ms))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record()
        gpu_pyramid()
        ev[s][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- e2e through the reference-facing API: host numpy in (pinned), host int64 out
    NN = S.nearest_neighbors
    pinned = _lib.pinned_empty(host_clouds.shape, np.float32)
    pinned[...] = host_clouds

    def api_pyramid():
        xyz = pinned
        h2d = d2h = 0
        for ratio in RATIOS:
            idx = NN.knn_batch(xyz, xyz, K, omp=True)
            h2d += xyz.nbytes
            d2h += idx.nbytes
            sub = xyz[:, : xyz.shape[1] // ratio, :]
            up = NN.knn_batch(sub, xyz, 1, omp=True)
            h2d += sub.nbytes + xyz.nbytes
            d2h += up.nbytes
            xyz = sub
        return h2d, d2h

    for _ in range(2):
        h2d_b, d2h_b = api_pyramid()
    e2e_steps = max(3, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        api_pyramid()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary numbers of the same hot path (rank 0, short): grid subsampling and FPS, device resident
    extra = {}
    if rank == 0 and not args.no_extra:
        extra = secondary_metrics(torch, D, dev, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    q0 = B * N0
    algo_bytes = q0 * (12 + 8 * K) + q0 * 12  # queries + int64 rows + support points (140 B/query + 12 B/pt)
    achieved = algo_bytes / (dom_ms_avg * 1e-3) / 1e9
    tie_rows = sum(st["tie_rows"] for _, _, st in stats)
    evals0 = stats[0][2]["dist_evals"]
    line = {
        "metric": METRIC, "value": K16_QUERIES * world * args.steps / (total_ms * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "points": N0, "k": K, "levels": LEVELS,
                   "l2": "flushed between timed iterations (256 MiB write)", "timing": "CUDA events per step, summed",
                   "k1_queries_per_step": K1_QUERIES, "k16_queries_per_step": K16_QUERIES},
        "e2e": {"value": K16_QUERIES * world * e2e_steps / e2e_dt, "unit": "queries/s",
                "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b), "steps": e2e_steps,
                "api": "ssdr_al_b200.nearest_neighbors.knn_batch (numpy in/out, int64 indices)"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("knn_query_kernel_level0"), "peak_source": peak_src, "kernel": "knn::query_kernel<17,int64> level 0 (6x40960)",
                     "kernel_ms": dom_ms_avg, "algorithmic_bytes": algo_bytes,
                     "note": "KNN is FP32/issue bound, not HBM bound (SURVEY.md 8d); see dist_evals_per_s",
                     "dist_evals_per_s": evals0 / (dom_ms_avg * 1e-3), "dist_evals_per_query": evals0 / q0},
        "clocks": clocks,
        "knn_detail": {"tie_rows_per_step": int(tie_rows), "stage_ms": [
            {"call": kind, "level_points": LEVELS[li], "grid_build_ms": st["grid_build_ms"],
             "main_kernel_ms": st["main_kernel_ms"], "tie_path_ms": st["tie_path_ms"], "tree_build_ms": st["tree_build_ms"],
             "tie_rows": st["tie_rows"]}
            for kind, li, st in stats]},
        "extra": extra,
    }
    # CPU baseline (bounded sample: 2 pyramid steps after 1 warm-up)
    try:
        fn, kind, cores = cpu_reference_fn()
        cpu_pyramid(host_clouds, fn)
        t0 = time.perf_counter()
        for _ in range(2):
            cpu_pyramid(host_clouds, fn)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": K16_QUERIES * 2 / dt, "unit": "queries/s", "cores": cores, "kind": kind,
                                "sample": "2 full pyramid steps (same 6x40960 clouds), knn_batch(omp=True), host has "
                                          "%d cpus" % (os.cpu_count() or 0)}
    except Exception as e:  # the oracle is test infrastructure; its absence must not hide the GPU number
        line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "port", "sample": "failed: %r" % e}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def secondary_metrics(torch, D, dev, flush):
    """Grid subsampling (config-1 shape) and FPS / k-center (config-4 shape), device resident, CUDA-event timed."""
    out = {}
    peak, _ = measured_peak()

    def timed(fn, reps):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), r

    def cpu_time(fn):
        t0 = time.perf_counter()
        fn()
        return (time.perf_counter() - t0) * 1e3

    try:
        from oracle import oracle as O
        O.lib()
    except Exception:
        O = None

    try:
        import ssdr_al_b200 as S
        rng = np.random.default_rng(0)
        n = 1_000_000
        face = rng.integers(0, 3, n)
        p = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
        p[face == 0, 2] = 0
        p[face == 1, 1] = 0
        p[face == 2, 0] = 0
        p += rng.normal(0, 0.005, p.shape)
        p -= p.min(0)
        p = p.astype(np.float32)
        rgb_h = rng.integers(0, 256, (n, 3)).astype(np.uint8)
        lab_h = ((p[:, 0] * 1.7).astype(np.int32) % 13).astype(np.uint8)
        pts = torch.from_numpy(p).to(dev)
        rgb = torch.from_numpy(rgb_h.astype(np.float32)).to(dev)
        lab = torch.from_numpy(lab_h.astype(np.int32)).to(dev)
        D.grid_subsample(pts, rgb, lab, 0.04)
        ms, r = timed(lambda: D.grid_subsample(pts, rgb, lab, 0.04), 5)
        m = r[0].shape[0]
        algo = n * 28 + m * 28
        g = {"points": n, "voxels": int(m), "ms": ms, "mpts_per_s": n / ms / 1e3,
             "algorithmic_gbs": algo / ms / 1e6, "frac_of_hbm_peak": algo / ms / 1e6 / peak}
        # the same call through the reference-facing API (uint8 colours / labels in, conversions + PCIe inside)
        S.grid_subsampling.compute(p, features=rgb_h, classes=lab_h, sampleDl=0.04)
        g["e2e_ms"] = float(np.median([cpu_time(lambda: S.grid_subsampling.compute(
            p, features=rgb_h, classes=lab_h, sampleDl=0.04)) for _ in range(3)]))
        g["e2e_mpts_per_s"] = n / g["e2e_ms"] / 1e3
        if O is not None:
            ref = O.ref_grid_subsample if O.have_ref() else O.grid_subsample
            g["cpu_ms"] = cpu_time(lambda: ref(p, rgb_h.astype(np.float32), lab_h.astype(np.int32), 0.04))
            g["cpu_kind"] = "reference (1 thread, one run)" if O.have_ref() else "port (1 thread, one run)"
        out["grid_subsample"] = g
        # config 1, second half: k=16 KNN of the sub-sampled cloud on itself (one cloud, N = Q = M)
        sub = r[0].contiguous()[None]
        sub2 = (sub * 1.5).contiguous()  # a second cloud of the same shape: alternating them keeps the library from
        clouds2 = [sub, sub2]            # reusing the previous call's trees, which a real caller would not get
        D.knn_batch(sub, sub, K)
        turn = [0]

        def knn_once():
            turn[0] ^= 1
            c = clouds2[turn[0]]
            return D.knn_batch(c, c, K)

        ms, _ = timed(knn_once, 6)
        out["knn_cfg1"] = {"points": int(m), "k": K, "ms": ms, "queries_per_s": m / ms * 1e3}
        # worst case for the subsampler: uniform in volume, M ~ 0.88 N
        u = torch.from_numpy((rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])).astype(np.float32)).to(dev)
        D.grid_subsample(u, rgb, lab, 0.04)
        ms, r = timed(lambda: D.grid_subsample(u, rgb, lab, 0.04), 5)
        out["grid_subsample_uniform"] = {"points": n, "voxels": int(r[0].shape[0]), "ms": ms,
                                         "mpts_per_s": n / ms / 1e3}
        del pts, rgb, lab, u, sub, sub2
    except Exception as e:
        out["grid_subsample"] = {"error": repr(e)}
    try:  # the same pyramid on SURFACE crops: what the real pipeline feeds (40960 nearest points of a random centre
        # in a room made of planes, shuffled) -- the headline config is uniform-in-volume by definition
        rng = np.random.default_rng(5)
        n = 400_000
        kk = rng.integers(0, 8, n)
        room = rng.random((n, 3)) * np.array([7.0, 5.0, 3.0])
        for v, (ax, val) in enumerate(((2, 0.0), (2, 3.0), (1, 0.0), (1, 5.0), (0, 0.0), (0, 7.0))):
            room[kk == v, ax] = val
        room[kk >= 6, 2] = 0.75
        room += rng.normal(0, 0.004, room.shape)
        room = room.astype(np.float32)
        crops = []
        for _ in range(B):
            dd = ((room - room[rng.integers(0, n)]) ** 2).sum(1)
            crops.append(room[rng.permutation(np.argpartition(dd, N0)[:N0])])
        cx = torch.from_numpy(np.stack(crops)).to(dev)

        def crop_pyramid():
            xyz = cx
            for ratio in RATIOS:
                D.knn_batch(xyz, xyz, K)
                sub = xyz[:, : xyz.shape[1] // ratio, :].contiguous()
                D.knn_batch(sub, xyz, 1)
                xyz = sub

        crop_pyramid()
        ms, _ = timed(crop_pyramid, 7)
        out["pyramid_surface_crops"] = {"batch": B, "points": N0, "ms": ms, "k16_queries_per_s": K16_QUERIES / ms * 1e3}
        del cx
    except Exception as e:
        out["pyramid_surface_crops"] = {"error": repr(e)}
    try:  # chamfer adjacency of one room's superpoints (fps_gcn_cpu.py:25-38), host arrays in / matrix out
        import ssdr_al_b200 as S
        rng = np.random.default_rng(9)
        sizes = rng.integers(60, 700, 300)
        sps, cents = [], []
        for nn_ in sizes:
            cc = rng.random(3) * np.array([7.0, 5.0, 3.0])
            pp = (cc + rng.normal(0, 0.2, (int(nn_), 3)) * rng.choice([1.0, 0.05], 3)).astype(np.float32)
            sps.append(pp)
            cents.append((pp.min(0).astype(np.float64) + pp.max(0)) / 2.0)
        cents = np.array(cents)
        S.chamfer.create_cd(sps, cents)
        ms = float(np.median([cpu_time(lambda: S.chamfer.create_cd(sps, cents)) for _ in range(3)]))
        tot = int(sizes.sum())
        out["chamfer_adjacency"] = {"superpoints": len(sps), "points": tot, "e2e_ms": ms,
                                    "pair_evals_per_s": float(tot) * tot / ms * 1e3}
        S.chamfer.farthest_superpoint_sample(sps, cents, 8, 0)
        n_pick = 60
        ms_f = float(np.median([cpu_time(lambda: S.chamfer.farthest_superpoint_sample(sps, cents, n_pick, 0))
                                for _ in range(3)]))
        out["superpoint_fps"] = {"superpoints": len(sps), "picks": n_pick, "e2e_ms": ms_f,
                                 "ms_per_pick": ms_f / (n_pick - 1)}
        if O is not None:  # the reference's KD-tree loop on the first 24 superpoints, scaled by the pair count
            sub = 24
            t_ref = cpu_time(lambda: (O.ref_create_cd if os.path.isdir("/root/reference") else O.create_cd)(
                sps[:sub], cents[:sub]))
            out["chamfer_adjacency"]["cpu_ms_extrapolated"] = t_ref * (len(sps) * (len(sps) - 1)) / (sub * (sub - 1))
            out["chamfer_adjacency"]["cpu_kind"] = "create_cd on %d superpoints, scaled by pairs (%s)" % (
                sub, "reference KD trees" if os.path.isdir("/root/reference") else "numpy restatement")
    except Exception as e:
        out["chamfer_adjacency"] = {"error": repr(e)}
    for d_, picks in ((32, 2000), (256, 1000)):
        try:
            g = torch.Generator(device=dev)
            g.manual_seed(3)
            F = torch.randn((500_000, d_), generator=g, device=dev, dtype=torch.float32)
            D.fps(F, 64, 12345)
            ms, _ = timed(lambda: D.fps(F, picks, 12345), 3)
            per = ms / (picks - 1)
            algo = 500_000 * (4 * d_ + 8)
            out["fps_d%d" % d_] = {"rows": 500_000, "picks": picks, "ms_per_pick": per, "picks_per_s": 1e3 / per,
                                   "algorithmic_gbs": algo / per / 1e6, "frac_of_hbm_peak": algo / per / 1e6 / peak}
            sel = torch.arange(500_000 - 16, 500_000, device=dev, dtype=torch.int64)
            D.kcenter(F, sel, 16)
            ms, _ = timed(lambda: D.kcenter(F, sel, picks), 3)
            per = ms / (picks + 16 - 1)
            algo = 500_000 * (4 * d_ + 8 + 8)
            out["kcenter_d%d" % d_] = {"rows": 500_000, "picks": picks, "ms_per_pick": per, "picks_per_s": 1e3 / per,
                                       "algorithmic_gbs": algo / per / 1e6, "frac_of_hbm_peak": algo / per / 1e6 / peak}
            if O is not None:  # the reference loops on the host, a few picks, extrapolated linearly
                Fh = F.cpu().numpy()
                npk = 8 if d_ == 32 else 4
                out["fps_d%d" % d_]["cpu_ms_per_pick"] = cpu_time(lambda: O.fps_numpy(Fh, npk + 1, 12345)) / npk
                out["fps_d%d" % d_]["cpu_kind"] = "numpy loop of fps_gcn_cpu.py:137-146, %d picks, 1 process" % npk
                selh = np.arange(500_000 - 4, 500_000)
                out["kcenter_d%d" % d_]["cpu_ms_per_pick"] = cpu_time(lambda: O.kcenter(Fh, selh, npk)) / (npk + 4)
                out["kcenter_d%d" % d_]["cpu_kind"] = "sklearn-formula restatement, %d centres, host BLAS threads" % (npk + 4)
                del Fh
            del F
        except Exception as e:
            out["fps_d%d" % d_] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary grid/FPS measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
