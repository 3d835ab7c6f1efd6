#!/usr/bin/env python
"""Turns gpurun_out/launches.csv + gpurun_out/prof_*.ncu-rep into the tracked summaries under profiles/:
   profiles/rNN_launches.txt      per-kernel share of one UNet forward + VAE decode
   profiles/rNN_ncu_<name>.csv    selected raw metrics per captured launch (DRAM bytes, tensor-pipe %, ...)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches():
    p = os.path.join(SRC, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("gyre::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    with open(os.path.join(OUT, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: one SD1.5 UNet forward (CFG batch 16, 64x64 "
                f"latents) + one VAE decode (batch 8 -> 512x512)\n# {n} launches, {tot / 1e3:.2f} ms summed device time "
                f"(cold-cache, serialised under the profiler: compare shares, not absolutes)\n")
        f.write(f"{'us':>12} {'share':>7} {'count':>6} {'avg us':>10}  kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:12.1f} {100 * t / tot:6.1f}% {c:6d} {t / c:10.1f}  {k}\n")
    print("wrote", f"{tag}_launches.txt")


def ncu_reports():
    for fn in sorted(os.listdir(SRC)):
        if not fn.endswith(".ncu-rep"):
            continue
        r = subprocess.run(["ncu", "-i", os.path.join(SRC, fn), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
        stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
        name = fn[:-8]
        with open(os.path.join(OUT, f"{tag}_ncu_{name}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols] +
                       [hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "stall_") for i in stall])
            for row in rows[2:]:
                w.writerow([re.sub(r"\(.*", "", row[cols[0]])] + [row[i] for i in cols[1:]] + [row[i] for i in stall])
        print("wrote", f"{tag}_ncu_{name}.csv", len(rows) - 2, "launches")


launches()
ncu_reports()
