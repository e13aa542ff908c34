#!/usr/bin/env python
"""bench.py -- BASELINE config 2: the RandLA-Net S3DIS input pyramid (knn_batch k=16 for 6 x 40,960 points plus the
four random-downsampled layers 10240/2560/640/160, and the 1-NN up-sampling query at every level), i.e. the loop of
SSDR_AL_s3dis/s3dis_dataset.py:164-177.

One "step" = one full pyramid for one batch of 6 synthetic clouds (seeded uniform box).
  value  = k=16 KNN queries/s with the clouds already resident in HBM (device-pointer C ABI, CUDA-event timed; the
           1-NN up-sampling work is inside the timed region but only k=16 queries are counted).
  e2e    = the same metric through the reference-facing Python API (`nearest_neighbors.knn_batch`, pageable host
           numpy in, host int64 out), host<->device copies inside the timed region.
  roofline     = the dominant kernel (level-0 k=16 query kernel) against the measured HBM copy peak, with the
                 issue-side numbers that actually bound it.
  cpu_baseline = the reference's own nanoflann/OpenMP code (oracle/_ref, or the C port when _ref is absent) timed on
                 this box's host cores on the same pyramid, OpenMP team pinned explicitly.
`--impl reference` runs only that CPU reference, as the driver's reference arm.

Multi-GPU (torchrun, one rank per GPU):
  * headline: batch items are independent, every rank runs its own batch of 6 clouds, no data-path collective (weak).
  * `extra.multi_gpu` (strong scaling, total work fixed as N grows; the same section runs at N = 1 so the driver's
    N = 1, 2, 4, 8 lines are comparable): config 3 (one Semantic3D-scale scan: slab-sharded subsampling -> all-gather
    of the slabs -> query-sharded k=16 KNN), config 4 (row-sharded FPS / k-center, the per-pick exchange fused into the
    persistent kernel over NVLink peer memory), config 5 (272 rooms round-robin + one global selection).
  * at N > 1 every sharded result is compared bit for bit with the single-GPU result IN THIS PROCESS
    (tools/multigpu_check.py + the full-size checks below); a mismatch is printed in the line and the exit code is 1.
"""
import argparse
import contextlib
import io
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
# identical in both arms (the driver compares the dicts)
CONFIG = {"workload": WORKLOAD, "batch_per_gpu": B, "points": N0, "k": K, "levels": LEVELS,
          "l2": "flushed between timed iterations (256 MiB write)", "timing": "CUDA events per step, summed",
          "k1_queries_per_step": K1_QUERIES, "k16_queries_per_step": K16_QUERIES}
CPU_THREADS = min(B, os.cpu_count() or 1)  # knn_batch(omp=True) parallelises over the 6 batch items (knn_.cxx:108-109)


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


def ncu_issue(kernel):
    """Issue-side numbers of the same capture (issue slots busy, warp instructions per query, lanes active)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)[kernel]
        return {k: d[k] for k in ("issue_active_pct", "warp_inst_per_query", "lanes_active", "fma_pipe_pct") if k in d}
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
    """(knn_batch callable, kind, OpenMP threads).  torchrun exports OMP_NUM_THREADS=1 to every rank; the team size is
    therefore set explicitly (and reported) instead of inherited."""
    from oracle import oracle as O
    t = O.set_omp_threads(CPU_THREADS)
    if O.have_ref():
        return (lambda p, q, k: O.ref_knn_batch(p, q, k, omp=True)), "reference", t
    O.lib()
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
        "config": CONFIG,
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": kind,
                         "omp_num_threads": cores, "host_cpus": os.cpu_count(),
                         "sample": "%d full pyramid steps, knn_batch(omp=True): OpenMP over the %d batch items, one "
                                   "host process (rank 0) whatever --gpus says" % (args.steps, B)},
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

    def gpu_pyramid(collect=None, xyz=None):
        xyz = xyz0 if xyz is None else xyz
        for li, ratio in enumerate(RATIOS):
            r = D.knn_batch(xyz, xyz, K, out=outs16[li], want_stats=collect is not None)
            if collect is not None:
                collect.append(("k16", li, r[1]))
            sub = xyz[:, : xyz.shape[1] // ratio, :].contiguous()
            r1 = D.knn_batch(sub, xyz, 1, out=outs1[li], want_stats=collect is not None)
            if collect is not None:
                collect.append(("k1", li, r1[1]))
            xyz = sub

    def count_launches(stats):
        # counted inside the library where the launches happen (memsets and torch's slicing copies not included)
        return sum(int(st["kernel_launches"]) for _, _, st in stats)

    def gpu_pyramid_fused(xyz=None):
        # the same ten queries through ONE C-ABI call (ssdr_knn_pyramid_dev): no host round trip between the levels,
        # the support clouds searched side by side
        D.knn_pyramid(xyz0 if xyz is None else xyz, RATIOS, K, neigh=outs16, up=outs1)

    headline = gpu_pyramid if args.per_call else gpu_pyramid_fused
    for _ in range(max(args.warmup, 3)):
        gpu_pyramid()
        headline()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().ssdr_knn_status(None))
    # the fused call must return the rows of the ten separate calls, bit for bit
    want16, want1 = [o.clone() for o in outs16], [o.clone() for o in outs1]
    gpu_pyramid()
    torch.cuda.synchronize()
    pyramid_equal = all(torch.equal(a, b_) for a, b_ in zip(want16 + want1, outs16 + outs1))

    # one instrumented pass (untimed) for per-kernel timing, tie statistics and launch counts
    stats = []
    flush.zero_()
    gpu_pyramid(stats)
    torch.cuda.synchronize()
    launches_per_step = count_launches(stats)
    if not args.per_call:
        gpu_pyramid_fused()
        torch.cuda.synchronize()
        launches_per_step = int(_lib.lib().ssdr_knn_pyramid_launches())
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
        headline()
        ev[s][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- e2e through the reference-facing API: host numpy in, host int64 out.  The caller's arrays are ordinary
    # (pageable) numpy memory, like the dataset loader's (s3dis_dataset.py:152-166); a second figure with the input
    # placed in pinned memory is kept beside it.
    NN = S.nearest_neighbors
    pinned = _lib.pinned_empty(host_clouds.shape, np.float32)
    pinned[...] = host_clouds

    def api_pyramid(xyz):
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

    def time_api(src):
        for _ in range(2):
            hb, db = api_pyramid(src)
        steps = max(3, min(args.steps, 10))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            api_pyramid(src)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return K16_QUERIES * world * steps / dt, hb, db, steps

    e2e_val, h2d_b, d2h_b, e2e_steps = time_api(host_clouds)
    e2e_pinned, _, _, _ = time_api(pinned)

    # the same pyramid through the ONE-call host entry (nearest_neighbors.knn_pyramid = the loop of
    # s3dis_dataset.py:164-177 as a single C-ABI call): one upload, rows copied back under the other levels' kernels
    def api_pyramid_one_call(xyz):
        neigh, up = NN.knn_pyramid(xyz, RATIOS, K)
        return xyz.nbytes, sum(a.nbytes for a in neigh) + sum(a.nbytes for a in up)

    def time_one_call(src):
        for _ in range(2):
            hb, db = api_pyramid_one_call(src)
        steps = max(3, min(args.steps, 10))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            api_pyramid_one_call(src)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return K16_QUERIES * world * steps / dt, hb, db

    one_val, one_h2d, one_d2h = time_one_call(host_clouds)
    got_n, got_u = NN.knn_pyramid(host_clouds, RATIOS, K)
    one_equal = all(np.array_equal(a, b_.cpu().numpy()) for a, b_ in zip(got_n + got_u, want16 + want1))
    del got_n, got_u
    clocks = sampler.stop() if rank == 0 else None

    # ---- secondary numbers of the same hot path (rank 0, short): grid subsampling, config-1 KNN, FPS / k-center
    extra = {}
    if rank == 0 and not args.no_extra:
        extra = secondary_metrics(torch, D, dev, flush, gpu_pyramid, gpu_pyramid_fused)
    # ---- sharded paths: all ranks take part (at N = 1 the same workloads run on the one GPU)
    mg_ok = True
    if not args.no_extra and not args.no_multi:
        try:
            mg, mg_ok = multi_gpu_metrics(torch, dist, D, dev, rank, world, flush)
        except Exception as e:  # a failure here must not hide the headline: it is recorded in the line
            mg, mg_ok = {"error": repr(e)}, True
        if world > 1:
            flag = torch.tensor([1 if mg_ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            mg_ok = int(flag.item()) == 1
        extra["multi_gpu"] = mg

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        sys.exit(0 if mg_ok else 1)

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
        "config": CONFIG,
        "e2e": {"value": e2e_val, "unit": "queries/s",
                "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b), "steps": e2e_steps,
                "input": "pageable numpy arrays (what the dataset loader passes)", "pinned_input_value": e2e_pinned,
                "api": "ssdr_al_b200.nearest_neighbors.knn_batch (numpy in/out, int64 indices), ten calls per pyramid "
                       "exactly like the reference loop",
                "one_call": {"value": one_val, "unit": "queries/s", "h2d_bytes_per_step": int(one_h2d),
                             "d2h_bytes_per_step": int(one_d2h), "equals_ten_calls": bool(one_equal),
                             "api": "ssdr_al_b200.nearest_neighbors.knn_pyramid (the same loop as ONE C-ABI call, "
                                    "ssdr_knn_pyramid; pageable numpy in, int64 numpy out)"}},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("knn_query_kernel_level0"), "peak_source": peak_src,
                     "kernel": "knn::query16_kernel<int64> level 0 (6x40960)",
                     "kernel_ms": dom_ms_avg, "algorithmic_bytes": algo_bytes,
                     "note": "KNN is FP32/issue bound, not HBM bound (SURVEY.md 8d): the numbers that bound it are "
                             "dist_evals_per_s and `issue` (ncu capture of the same launch)",
                     "dist_evals_per_s": evals0 / (dom_ms_avg * 1e-3), "dist_evals_per_query": evals0 / q0,
                     "issue": ncu_issue("knn_query_kernel_level0")},
        "clocks": clocks,
        "headline_call": ("ten ssdr_knn_batch_dev calls" if args.per_call else
                          "one ssdr_knn_pyramid_dev call (no host round trip; the support clouds run as concurrent branches, "
                          "replayed as one CUDA graph)"),
        "pyramid_call_equals_per_call_results": bool(pyramid_equal),
        "knn_detail": {"tie_rows_per_step": int(tie_rows), "stage_ms": [
            {"call": kind, "level_points": LEVELS[li], "grid_build_ms": st["grid_build_ms"],
             "main_kernel_ms": st["main_kernel_ms"], "tie_path_ms": st["tie_path_ms"], "tree_build_ms": st["tree_build_ms"],
             "tie_rows": st["tie_rows"]}
            for kind, li, st in stats]},
        "extra": extra,
    }
    # CPU baseline (bounded sample: 2 pyramid steps after 1 warm-up), team size pinned like the reference arm's
    try:
        fn, kind, cores = cpu_reference_fn()
        cpu_pyramid(host_clouds, fn)
        t0 = time.perf_counter()
        for _ in range(2):
            cpu_pyramid(host_clouds, fn)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": K16_QUERIES * 2 / dt, "unit": "queries/s", "cores": cores, "kind": kind,
                                "omp_num_threads": cores, "host_cpus": os.cpu_count(),
                                "sample": "2 full pyramid steps (same 6x40960 clouds), knn_batch(omp=True), OpenMP team "
                                          "set to %d explicitly, host has %d cpus" % (cores, os.cpu_count() or 0)}
    except Exception as e:  # the oracle is test infrastructure; its absence must not hide the GPU number
        line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "port", "sample": "failed: %r" % e}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if mg_ok else 1)


def _timed(torch, flush, fn, reps):
    ts = []
    r = None
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), r


def _cpu_ms(fn):
    t0 = time.perf_counter()
    fn()
    return (time.perf_counter() - t0) * 1e3


def _oracle():
    try:
        from oracle import oracle as O
        O.lib()
        return O
    except Exception:
        return None


def _grid_traffic():
    """DRAM bytes of one config-1 subsampling call: the ncu captures of its kernels (pack, sort, the two reduces)."""
    parts = [ncu_traffic(k) for k in ("grid_pack_1m", "grid_sort_1m", "grid_reduce_1m", "grid_reduce_groups_1m")]
    return None if any(v is None for v in parts) else float(sum(parts))


def secondary_metrics(torch, D, dev, flush, gpu_pyramid, gpu_pyramid_fused):
    """Grid subsampling + KNN (config-1 shape), FPS / k-center (config-4 shape), the tie-heavy pyramid (8f-1), chamfer
    adjacency: device resident, CUDA-event timed, each with its end-to-end and CPU-reference figure beside it."""
    from tools import synth
    out = {}
    peak, _ = measured_peak()
    O = _oracle()
    timed = lambda fn, reps: _timed(torch, flush, fn, reps)  # noqa: E731

    try:
        import ssdr_al_b200 as S
        n = 1_000_000
        p, rgb_h, lab_h = synth.room_cloud(n, 0)  # floor, ceiling, walls, furniture; ~9 points / voxel
        pts = torch.from_numpy(p).to(dev)
        rgb = torch.from_numpy(rgb_h.astype(np.float32)).to(dev)
        lab = torch.from_numpy(lab_h.astype(np.int32)).to(dev)
        D.grid_subsample(pts, rgb, lab, 0.04)
        ms, r = timed(lambda: D.grid_subsample(pts, rgb, lab, 0.04), 7)
        m = r[0].shape[0]
        algo = n * 28 + m * 28
        g = {"points": n, "voxels": int(m), "ms": ms, "mpts_per_s": n / ms / 1e3,
             "algorithmic_gbs": algo / ms / 1e6, "frac_of_hbm_peak": algo / ms / 1e6 / peak,
             "traffic_bytes": _grid_traffic()}
        # the same call through the reference-facing API (uint8 colours / labels in, conversions + PCIe inside)
        S.grid_subsampling.compute(p, features=rgb_h, classes=lab_h, sampleDl=0.04)
        g["e2e_ms"] = float(np.median([_cpu_ms(lambda: S.grid_subsampling.compute(
            p, features=rgb_h, classes=lab_h, sampleDl=0.04)) for _ in range(3)]))
        g["e2e_mpts_per_s"] = n / g["e2e_ms"] / 1e3
        if O is not None:
            ref = O.ref_grid_subsample if O.have_ref() else O.grid_subsample
            g["cpu_ms"] = _cpu_ms(lambda: ref(p, rgb_h.astype(np.float32), lab_h.astype(np.int32), 0.04))
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
        kc = {"points": int(m), "k": K, "ms": ms, "queries_per_s": m / ms * 1e3}
        sub_h = [c[0].cpu().numpy() for c in clouds2]
        S.nearest_neighbors.knn(sub_h[0], sub_h[0], K, omp=True)

        def knn_api():
            turn[0] ^= 1
            return S.nearest_neighbors.knn(sub_h[turn[0]], sub_h[turn[0]], K, omp=True)

        kc["e2e_ms"] = float(np.median([_cpu_ms(knn_api) for _ in range(4)]))
        kc["e2e_queries_per_s"] = m / kc["e2e_ms"] * 1e3
        if O is not None:
            threads = O.set_omp_threads(os.cpu_count() or 1)  # cpp_knn_omp parallelises over the queries
            fn = (lambda: O.ref_knn(sub_h[0], sub_h[0], K, omp=True)) if O.have_ref() else (
                lambda: O.knn(sub_h[0], sub_h[0], K, threads=threads))
            kc["cpu_ms"] = _cpu_ms(fn)
            kc["cpu_kind"] = "%s knn(omp=True), %d OpenMP threads, one run" % (
                "reference" if O.have_ref() else "port", threads)
            O.set_omp_threads(CPU_THREADS)
        out["knn_cfg1"] = kc
        # worst case for the subsampler: uniform in volume, M ~ 0.88 N
        rng = np.random.default_rng(0)
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
        room, _, _ = synth.room_cloud(400_000, 5, noise=0.004)
        crops = []
        for _ in range(B):
            dd = ((room - room[rng.integers(0, len(room))]) ** 2).sum(1)
            crops.append(room[rng.permutation(np.argpartition(dd, N0)[:N0])])
        crops = np.stack(crops)
        cx = torch.from_numpy(crops).to(dev)
        for _ in range(3):
            gpu_pyramid_fused(cx)
        ms, _ = timed(lambda: gpu_pyramid_fused(cx), 7)
        out["pyramid_surface_crops"] = {"batch": B, "points": N0, "ms": ms, "k16_queries_per_s": K16_QUERIES / ms * 1e3}
        # duplicate-heavy input -- the loader's data_aug pads short crops by REPEATING points
        # (s3dis_dataset.py:147-150), so most rows of such a batch hold exact distance ties and take the tie path
        dup = crops.copy()
        for b in range(B):
            keep = N0 // 2
            dup[b, keep:] = dup[b, rng.integers(0, keep, N0 - keep)]
            dup[b] = dup[b, rng.permutation(N0)]
        cd = torch.from_numpy(dup).to(dev)
        st = []
        gpu_pyramid(collect=st, xyz=cd)
        for _ in range(3):
            gpu_pyramid_fused(cd)
        ms, _ = timed(lambda: gpu_pyramid_fused(cd), 5)
        out["pyramid_duplicated_points"] = {
            "batch": B, "points": N0, "duplicated_fraction": 0.5, "ms": ms, "k16_queries_per_s": K16_QUERIES / ms * 1e3,
            "tie_rows_per_step": int(sum(s_[2]["tie_rows"] for s_ in st)),
            "tie_path_ms": float(sum(s_[2]["tie_path_ms"] for s_ in st))}
        if O is not None:
            fn, kind, cores = cpu_reference_fn()
            out["pyramid_duplicated_points"]["cpu_ms"] = _cpu_ms(lambda: cpu_pyramid(dup, fn))
            out["pyramid_duplicated_points"]["cpu_kind"] = "%s, %d threads" % (kind, cores)
        del cx, cd
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
        ms = float(np.median([_cpu_ms(lambda: S.chamfer.create_cd(sps, cents)) for _ in range(3)]))
        tot = int(sizes.sum())
        out["chamfer_adjacency"] = {"superpoints": len(sps), "points": tot, "e2e_ms": ms,
                                    "pair_evals_per_s": float(tot) * tot / ms * 1e3}
        S.chamfer.farthest_superpoint_sample(sps, cents, 8, 0)
        n_pick = 60
        ms_f = float(np.median([_cpu_ms(lambda: S.chamfer.farthest_superpoint_sample(sps, cents, n_pick, 0))
                                for _ in range(3)]))
        out["superpoint_fps"] = {"superpoints": len(sps), "picks": n_pick, "e2e_ms": ms_f,
                                 "ms_per_pick": ms_f / (n_pick - 1)}
        if O is not None:  # the reference's KD-tree loop on the first 24 superpoints, scaled by the pair count
            sub = 24
            have = O.have_ref_py()
            t_ref = _cpu_ms(lambda: (O.ref_py("fps_gcn_cpu").create_cd if have else O.create_cd)(sps[:sub], cents[:sub]))
            out["chamfer_adjacency"]["cpu_ms_extrapolated"] = t_ref * (len(sps) * (len(sps) - 1)) / (sub * (sub - 1))
            out["chamfer_adjacency"]["cpu_kind"] = "create_cd on %d superpoints, scaled by pairs (%s)" % (
                sub, "reference KD trees" if have else "numpy restatement")
    except Exception as e:
        out["chamfer_adjacency"] = {"error": repr(e)}
    try:  # adjacency + propagation in front of the FPS loop (fps_adj_all :95-116, GCN_FPS_sampling :153-167):
        # 8192 superpoints in 256 rooms, 32 features, one product; host arrays in and out like the reference
        import ssdr_al_b200 as S
        rng = np.random.default_rng(9)
        n_sp, per_room, d_feat = 8192, 32, 32
        perm = rng.permutation(n_sp)
        rooms = []
        for r0 in range(0, n_sp, per_room):
            cd = rng.random((per_room, per_room)) * 0.8
            cd = cd + cd.T
            np.fill_diagonal(cd, 0.0)
            rooms.append((perm[r0:r0 + per_room].tolist(), rng.random((per_room, 3)) * 8.0, cd))
        vfeat = rng.standard_normal((n_sp, d_feat))
        G = S.fps_gcn

        def gcn_once():
            a = G.adjacency_from_rooms(n_sp, rooms)
            r = G.propagate(a, vfeat, 1, 0)
            a.close()
            return r
        got = gcn_once()
        ms_g = float(np.median([_cpu_ms(gcn_once) for _ in range(3)]))
        out["gcn_adjacency_propagation"] = {
            "superpoints": n_sp, "rooms": len(rooms), "features": d_feat, "products": 1, "e2e_ms": ms_g,
            "matrix_bytes": 8 * n_sp * n_sp,
            "note": "host arrays in, host features out; the N x N matrix is assembled, normalised and multiplied on "
                    "the device and never leaves it"}
        if O is not None:
            t_ref = _cpu_ms(lambda: O.gcn_propagate(O.gcn_adjacency(n_sp, rooms), vfeat, 1, 0))
            want = O.gcn_propagate(O.gcn_adjacency(n_sp, rooms), vfeat, 1, 0)
            out["gcn_adjacency_propagation"]["cpu_ms"] = t_ref
            out["gcn_adjacency_propagation"]["cpu_kind"] = "numpy restatement of fps_gcn_cpu.py:63-116,153-167 (the "\
                                                           "reference IS these numpy calls), host BLAS threads"
            out["gcn_adjacency_propagation"]["max_rel_diff"] = float(np.max(np.abs(got - want) / (np.abs(want) + 1e-300)))
    except Exception as e:
        out["gcn_adjacency_propagation"] = {"error": repr(e)}
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
                                   "algorithmic_gbs": algo / per / 1e6, "frac_of_hbm_peak": algo / per / 1e6 / peak,
                                   "traffic_bytes_per_pick": ncu_traffic("fps_d%d" % d_)}
            sel = torch.arange(500_000 - 16, 500_000, device=dev, dtype=torch.int64)
            D.kcenter(F, sel, 16)
            ms, _ = timed(lambda: D.kcenter(F, sel, picks), 3)
            per = ms / (picks + 16 - 1)
            algo = 500_000 * (4 * d_ + 8 + 8)
            out["kcenter_d%d" % d_] = {"rows": 500_000, "picks": picks, "ms_per_pick": per, "picks_per_s": 1e3 / per,
                                       "algorithmic_gbs": algo / per / 1e6, "frac_of_hbm_peak": algo / per / 1e6 / peak}
            if O is not None:  # the reference's own loops on the host, a few picks, extrapolated linearly
                Fh = F.cpu().numpy()
                npk = 8 if d_ == 32 else 4
                selh = np.arange(500_000 - 4, 500_000)
                if O.have_ref_py():
                    fps_ref = O.ref_py("fps_gcn_cpu").farthest_features_sample
                    kcg = O.ref_py("kcenterGreedy").kCenterGreedy
                    out["fps_d%d" % d_]["cpu_ms_per_pick"] = _cpu_ms(lambda: fps_ref(Fh, npk + 1)) / npk
                    out["fps_d%d" % d_]["cpu_kind"] = ("reference farthest_features_sample (fps_gcn_cpu.py:119-147), "
                                                       "%d picks, 1 process" % npk)
                    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints progress lines
                        out["kcenter_d%d" % d_]["cpu_ms_per_pick"] = _cpu_ms(
                            lambda: kcg(Fh).select_batch_(selh, npk)) / (npk + 4)
                    out["kcenter_d%d" % d_]["cpu_kind"] = ("reference kCenterGreedy.select_batch_ (sklearn %s), %d "
                                                           "centres, host BLAS threads" % (_sklearn_version(), npk + 4))
                else:
                    out["fps_d%d" % d_]["cpu_ms_per_pick"] = _cpu_ms(lambda: O.fps_numpy(Fh, npk + 1, 12345)) / npk
                    out["fps_d%d" % d_]["cpu_kind"] = "port: numpy loop of fps_gcn_cpu.py:137-146, %d picks" % npk
                    out["kcenter_d%d" % d_]["cpu_ms_per_pick"] = _cpu_ms(lambda: O.kcenter(Fh, selh, npk)) / (npk + 4)
                    out["kcenter_d%d" % d_]["cpu_kind"] = "port: sklearn-formula restatement, %d centres" % (npk + 4)
                del Fh
            del F
        except Exception as e:
            out["fps_d%d" % d_] = {"error": repr(e)}
    return out


def _sklearn_version():
    try:
        import sklearn
        return sklearn.__version__
    except Exception:
        return "?"


def _max_ms(torch, dist, dev, world, ms):
    if world == 1:
        return float(ms)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def multi_gpu_metrics(torch, dist, D, dev, rank, world, flush):
    """Strong-scaling lines of the sharded paths (total work fixed; N = 1 runs the same work on one GPU), each checked
    bit for bit against the single-GPU result when N > 1.  Device time = CUDA events, max over ranks."""
    from ssdr_al_b200 import dist as SD
    from tools import synth
    out = {"world": world}
    ok = True
    comm = None
    peak, _ = measured_peak()
    if world > 1:
        from tools import multigpu_check as MC
        comm = MC.make_comm(dev)
        out["selection_transport"] = ("peer-memory mailboxes inside the persistent kernel (CUDA IPC over NVLink)"
                                      if isinstance(comm, SD.PeerGroup) else "nccl 8-byte all-reduce per pick (fallback)")
        chk_ok, chk = MC.run_checks(dev, rank, world, comm)
        out["equality_checks"] = {"all_equal": chk_ok, "results": chk}
        ok &= chk_ok

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def ev_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- config 4: 500k x {32, 256} float32, row-sharded FPS and k-center ------------------------------------------
    for d_, picks in ((32, 2000), (256, 1000)):
        try:
            g = torch.Generator(device=dev)
            g.manual_seed(3)  # the same matrix on every rank
            F = torch.randn((500_000, d_), generator=g, device=dev, dtype=torch.float32)
            sel = torch.arange(500_000 - 16, 500_000, device=dev, dtype=torch.int64)
            p2p = isinstance(comm, SD.PeerGroup)

            def run_fps():
                return D.fps(F, picks, 12345) if world == 1 else SD.fps_sharded(F, picks, 12345, comm)

            def run_kc():
                return D.kcenter(F, sel, picks) if world == 1 else SD.kcenter_sharded(F, sel, picks, comm)

            for name, fn, steps, extra_b in (("fps", run_fps, picks - 1, 8), ("kcenter", run_kc, picks + 15, 16)):
                if name == "kcenter" and world > 1 and not p2p:
                    continue
                fn()
                ts = []
                for _ in range(3):
                    barrier()
                    a, b = ev_pair()
                    a.record()
                    r = fn()
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(_max_ms(torch, dist, dev, world, a.elapsed_time(b)))
                per = float(np.median(ts)) / steps
                algo = 500_000 * (4 * d_ + extra_b)
                e = {"rows": 500_000, "picks": picks, "us_per_pick": 1e3 * per, "picks_per_s": 1e3 / per,
                     "algorithmic_gbs_all_gpus": algo / per / 1e6,
                     "frac_of_hbm_peak_per_gpu": algo / per / 1e6 / peak / world}
                if world > 1:
                    single = D.fps(F, picks, 12345) if name == "fps" else D.kcenter(F, sel, picks)
                    e["equal_to_single_gpu"] = bool(torch.equal(single, r))
                    ok &= e["equal_to_single_gpu"]
                    if p2p:  # per pick and rank: one 16-byte mailbox write to each peer (2 x 8-byte words, float32)
                        e["nvlink_bytes_per_pick_per_rank"] = 16 * (world - 1)
                out["%s_d%d" % (name, d_)] = e
            del F
        except Exception as e:  # recorded in the line; only a MISMATCH fails the run
            out["fps_d%d_error" % d_] = repr(e)

    # ---- config 3: one Semantic3D-scale scan.  subsample 0.06 (voxel-layer slabs) -> all-gather -> k=16 KNN ----------
    try:
        n_scan = int(os.environ.get("SSDR_BENCH_SCAN_POINTS", "80000000"))
        xyz, rgb, lab = synth.scan_cloud(n_scan, 2, dev)  # replicated on every rank (seeded device generator)
        lab2 = lab[:, None].contiguous()

        def scan_once():
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            evs[0].record()
            if world == 1:
                sp, sf, sc = D.grid_subsample(xyz, rgb, lab2, 0.06)
                evs[1].record()
                evs[2].record()
                idx = D.knn_batch(sp[None], sp[None], K)[0]
                full, span = sp, (0, sp.shape[0])
            else:
                sp, sf, sc = SD.grid_subsample_sharded(xyz, rgb, lab2, 0.06, replicated=True, axis="auto")
                evs[1].record()
                full, span = SD.gather_rows(sp)       # the exchange step: every rank needs the whole support cloud
                evs[2].record()
                idx = D.knn_batch(full[None], sp[None], K)[0] if sp.shape[0] else torch.zeros((0, K), dtype=torch.int64, device=dev)
            evs[3].record()
            torch.cuda.synchronize()
            t = [evs[i].elapsed_time(evs[i + 1]) for i in range(3)]
            return t, (sp, sf, sc, idx, full, span)

        scan_once()
        runs = []
        for _ in range(3):
            flush.zero_()
            barrier()
            t, res = scan_once()
            runs.append([_max_ms(torch, dist, dev, world, v) for v in t] + [_max_ms(torch, dist, dev, world, sum(t))])
        med = np.median(np.array(runs), axis=0)
        sp, sf, sc, idx, full, span = res
        m_total = int(full.shape[0])
        c3 = {"points": n_scan, "sampleDl": 0.06, "voxels": m_total, "subsample_ms": float(med[0]),
              "gather_ms": float(med[1]), "knn_ms": float(med[2]), "total_ms": float(med[3]),
              "subsample_mpts_per_s": n_scan / med[0] / 1e3, "knn_queries_per_s": m_total / med[2] * 1e3,
              "pipeline_mpts_per_s": n_scan / med[3] / 1e3,
              "subsample_algorithmic_gbs": (n_scan * 28 + m_total * 28) / med[0] / 1e6,
              "input": "replicated on every rank (seeded on-device generator); slabs = balanced voxel layers along "
                       "the best balanced axis (sampled layer histograms)",
              "note": "every component is the max over ranks, so a rank that waits for a slower one shows the wait in "
                      "gather_ms; the cell grid and the nanoflann-identical tree of the tie path are built on every rank "
                      "over the WHOLE cloud (replicated work): only the query scan of the KNN shards"}
        if world > 1:  # full-size check against this rank's own single-GPU run of the whole scan
            got = SD.grid_subsample_sharded(xyz, rgb, lab2, 0.06, replicated=True, axis="auto", return_keys=True,
                                            return_axis=True)
            c3["slab_axis"] = int(got[5])
            wp, wf, wc, wk, wn = D.grid_subsample(xyz, rgb, lab2, 0.06, return_keys=True)
            pos = np.searchsorted(wk, got[3])  # this rank's voxels inside the key-ordered single-GPU result
            tpos = torch.from_numpy(pos).to(dev)
            counts = torch.tensor([got[0].shape[0]], dtype=torch.int64, device=dev)
            dist.all_reduce(counts)
            same = (int(counts.item()) == wp.shape[0] == m_total and bool((pos < len(wk)).all())
                    and np.array_equal(wk[np.minimum(pos, len(wk) - 1)], got[3]) and np.array_equal(wn[pos], got[4])
                    and torch.equal(wp[tpos], got[0]) and torch.equal(wf[tpos], got[1]) and torch.equal(wc[tpos], got[2])
                    and torch.equal(got[0], sp))
            # KNN: the sharded rows against a single-GPU query of the SAME gathered cloud (nanoflann's tie order
            # depends on the order of the points, so the comparison keeps the order)
            widx = D.knn_batch(full[None], full[None], K)[0]
            b, e = span
            same = same and torch.equal(widx[b:e], idx)
            c3["equal_to_single_gpu"] = bool(same)
            ok &= bool(same)
            del wp, wf, wc, widx, got
        out["cfg3_scan"] = c3
        del xyz, rgb, lab, lab2, sp, sf, sc, idx, full, res
    except Exception as e:
        out["cfg3_scan"] = {"error": repr(e)}

    # ---- config 5: 272 rooms (per-room subsample 0.04 + k=16 KNN, rooms round-robin over the ranks), then ONE global
    # selection over ~500k superpoint feature rows (D = 32, 10,000 picks = 2 %) --------------------------------------
    try:
        n_rooms = int(os.environ.get("SSDR_BENCH_ROOMS", "272"))
        sizes = synth.room_sizes(n_rooms)
        mine = SD.shard_items(n_rooms, world, rank)
        room_ms = 0.0
        raw = vox = 0
        csum = torch.zeros(2, dtype=torch.int64, device=dev)
        for i in mine:
            p, f, c = synth.room_cloud_device(int(sizes[i]), 100 + i, dev)
            c2 = c[:, None].contiguous()
            a, b = ev_pair()
            a.record()
            sp, _, _ = D.grid_subsample(p, f, c2, 0.04)
            idx = D.knn_batch(sp[None], sp[None], K)
            b.record()
            torch.cuda.synchronize()
            room_ms += a.elapsed_time(b)
            raw += int(sizes[i])
            vox += int(sp.shape[0])
            csum[0] += idx.sum()
            csum[1] += sp.shape[0]
            del p, f, c, c2, sp, idx
        tot = torch.tensor([raw, vox], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
            dist.all_reduce(csum)
        room_ms = _max_ms(torch, dist, dev, world, room_ms)
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        F = torch.randn((500_000, 32), generator=g, device=dev, dtype=torch.float32)
        picks = 10_000
        fn = (lambda: D.fps(F, picks, 4242)) if world == 1 else (lambda: SD.fps_sharded(F, picks, 4242, comm))
        fn()
        barrier()
        a, b = ev_pair()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        sel_ms = _max_ms(torch, dist, dev, world, a.elapsed_time(b))
        c5 = {"rooms": n_rooms, "raw_points": int(tot[0]), "subsampled_points": int(tot[1]),
              "rooms_ms": room_ms, "rooms_per_s": n_rooms / room_ms * 1e3, "raw_mpts_per_s": int(tot[0]) / room_ms / 1e3,
              "knn_index_checksum": int(csum[0]), "selection_rows": 500_000, "selection_picks": picks,
              "selection_ms": sel_ms, "picks_per_s": picks / sel_ms * 1e3, "round_ms": room_ms + sel_ms}
        if world > 1:
            c5["selection_equal_to_single_gpu"] = bool(torch.equal(r, D.fps(F, picks, 4242)))
            ok &= c5["selection_equal_to_single_gpu"]
        out["cfg5_round"] = c5
    except Exception as e:
        out["cfg5_round"] = {"error": repr(e)}
    if comm is not None:
        barrier()
        comm.destroy()
    return out, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary grid/FPS measurements")
    ap.add_argument("--no-multi", action="store_true", help="skip the sharded-path section (configs 3, 4, 5)")
    ap.add_argument("--per-call", action="store_true", help="time the pyramid as ten knn_batch calls (round-1 headline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
