"""Summarise .ncu-rep captures (ncu --set full) and launch lists (gpu__time_duration) into small text files for
profiles/.  Usage: python tools/ncu_summary.py rep <file.ncu-rep> | launches <file.csv> |
traffic <file.ncu-rep> [key profiles/traffic.json]"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[idx["Kernel Name"]][:150])
        for k in KEYS:
            if k in idx:
                print("  %-62s %s %s" % (k, r[idx[k]], units[idx[k]]))
        stalls = []
        for h in hdr:
            if "issue_stalled" in h and "per_issue_active" in h:
                try:
                    v = float(r[idx[h]])
                except ValueError:
                    continue
                if v >= 0.2:
                    stalls.append((v, h.split("issue_stalled_")[1].split("_per_issue")[0]))
        print("  warp stall cycles per issued instruction (>=0.2):",
              ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)))
        print()


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(row["Kernel Name"][:100], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print("%-100s %6s %12s %10s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "max_us", "share"))
    for k, (n, t, m) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-100s %6d %12.1f %10.1f %10.1f %6.1f%%" % (k, n, t, t / n, m, 100 * t / tot))
    print("total_us %.1f  (ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)" % tot)


def traffic(path, key=None, out_json=None, units_per_launch=None):
    """dram bytes (read + write) per launch of the first kernel in an `ncu --set full` report; with key and out_json the
    value is merged into that json file (profiles/traffic.json, read by bench.py for roofline.traffic) together with the
    issue-side numbers of the same launch.  units_per_launch (queries, picks ...): per-unit figures are added."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[k]].replace(",", "")) * scale[units[idx[k]]]
    dur = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    print("%s: dram bytes per launch %.0f (%s %s under ncu)" % (r[idx["Kernel Name"]][:80], tot, dur,
                                                                 units[idx["gpu__time_duration.sum"]]))
    if key and out_json:
        try:
            d = json.load(open(out_json))
        except Exception:
            d = {}
        e = {"dram_bytes_per_launch": tot, "kernel": r[idx["Kernel Name"]][:120], "source": path.split("/")[-1]}

        def num(k):
            try:
                return float(r[idx[k]].replace(",", ""))
            except Exception:
                return None

        e["issue_active_pct"] = num("smsp__issue_active.avg.pct_of_peak_sustained_active")
        e["lanes_active"] = num("smsp__thread_inst_executed_per_inst_executed.ratio")
        e["fma_pipe_pct"] = num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
        e["l2_hit_pct"] = num("lts__t_sector_hit_rate.pct")
        inst = num("smsp__inst_executed.sum")
        if units_per_launch:
            u = float(units_per_launch)
            e["units_per_launch"] = u
            e["dram_bytes_per_unit"] = tot / u
            if inst is not None:
                e["warp_inst_per_query"] = inst / u
        d[key] = e
        json.dump(d, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    {"rep": rep, "launches": launches, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
