#!/usr/bin/env python
"""Turns gpurun_out/launches_{unet,vae}.csv + gpurun_out/prof_*.ncu-rep into the tracked summaries under profiles/:
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


def _read_launches(path):
    """[(kernel name, us, dram bytes)] per launch from an `ncu --metrics ... --csv` log (one row per metric)."""
    if not os.path.exists(path):
        return []
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        key = row["ID"]
        name = re.sub(r"\(CUtensorMap.*|\(.*", "", row["Kernel Name"]).replace("void ", "").replace("gyre::", "")
        rec = per.setdefault(key, [name, 0.0, 0.0])
        if row["Metric Name"] == "gpu__time_duration.sum":
            rec[1] = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            rec[2] += v * mult
    return list(per.values())


def launches():
    unet = _read_launches(os.path.join(SRC, "launches_unet.csv"))
    vae = _read_launches(os.path.join(SRC, "launches_vae.csv"))
    if not unet and not vae:
        return
    with open(os.path.join(OUT, f"{tag}_launches.txt"), "w") as f:
        for label, rows in (("one SD1.5 UNet forward (CFG batch 16, 64x64 latents)", unet),
                            ("one VAE decode (batch 8 -> 512x512)", vae)):
            agg = collections.OrderedDict()
            for name, us, by in rows:
                a = agg.setdefault(name, [0, 0.0, 0.0])
                a[0] += 1
                a[1] += us
                a[2] += by
            tot = sum(a[1] for a in agg.values()) or 1.0
            f.write(f"# ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none: {label}\n"
                    f"# {len(rows)} launches, {tot / 1e3:.2f} ms summed device time (cold-cache, serialised under the "
                    f"profiler: compare shares, not absolutes)\n")
            f.write(f"{'us':>12} {'share':>7} {'count':>6} {'avg us':>10} {'DRAM MB/launch':>15}  kernel\n")
            for k, (c, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{t:12.1f} {100 * t / tot:6.1f}% {c:6d} {t / c:10.1f} {by / c / 1e6:15.2f}  {k}\n")
            f.write("\n")
    # DRAM traffic per launch of each kernel family, weighted like one bench step (50 UNet forwards + 1 decode)
    fam = {}
    for rows, w in ((unet, 50), (vae, 1)):
        for name, us, by in rows:
            if "gemm_tc_kernel" in name:
                k = "conv3x3" if re.search(r"gemm_tc_kernel<\d+, 1[,>]", name) else "gemm"
            elif "attention" in name:
                k = "attention"
            elif name.startswith("gn_"):
                k = "groupnorm"
            elif "layernorm" in name:
                k = "layernorm"
            else:
                continue
            a = fam.setdefault(k, [0, 0.0, 0.0])
            a[0] += w
            a[1] += w * by
            a[2] += w * us
    import json
    json.dump({k: {"launches_per_step": v[0], "dram_bytes_per_launch": v[1] / v[0], "ncu_us_per_launch": v[2] / v[0]}
               for k, v in fam.items()}, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    print("wrote", f"{tag}_launches.txt", f"{tag}_traffic.json")


def ncu_reports():
    for fn in sorted(os.listdir(SRC)):
        if fn.endswith("_raw.csv") and fn.startswith("prof_"):      # raw page exported on the GPU box (gpu_profile.sh)
            rows = list(csv.reader(open(os.path.join(SRC, fn))))
            fn = fn[:-8] + ".ncu-rep"
        elif fn.endswith(".ncu-rep"):
            r = subprocess.run(["ncu", "-i", os.path.join(SRC, fn), "--page", "raw", "--csv"], capture_output=True, text=True)
            rows = list(csv.reader(io.StringIO(r.stdout)))
        else:
            continue
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
