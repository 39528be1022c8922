#!/usr/bin/env python
"""Turns the ncu launch lists / raw pages of scripts/ncu_all.sh into the short summaries kept under profiles/.
    python scripts/ncu_summarise.py gpurun_out/ncu_r2 profiles/r2g"""
import collections
import csv
import io
import os
import re
import sys

src, dst = sys.argv[1], sys.argv[2]


def launches(path):
    text = open(path).read()
    start = text.find('"ID"')
    per = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(io.StringIO(text[start:])):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"].lower()
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        n, t = per.get(name, (0, 0.0))
        per[name] = (n + 1, t + ns)
        total += ns
    return per, total


for f in sorted(os.listdir(src)):
    if f.startswith("launches_") and f.endswith(".csv"):
        per, total = launches(os.path.join(src, f))
        out = [f"# {f}: one transcription under ncu (gpu__time_duration.sum per launch, --clock-control none); shares, not bench values",
               f"{'kernel':42s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}"]
        for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            out.append(f"{name:42s} {n:8d} {t / 1e3:10.1f} {t / 1e3 / n:9.2f} {100 * t / total:6.1f}%")
        out.append(f"{'total':42s} {sum(n for n, _ in per.values()):8d} {total / 1e3:10.1f}")
        open(f"{dst}_{f[:-4]}_shares.txt", "w").write("\n".join(out) + "\n")
        print("\n".join(out[:14]))

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__warp_issue_stalled", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "smsp__average_warp",
        "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active", "gpc__cycles_elapsed.max", "sm__cycles_active.avg", "l1tex__data_bank_conflicts",
        "smsp__pcsamp_warps_issue_stalled"]
for f in sorted(os.listdir(src)):
    if not f.endswith(".raw.csv"):
        continue
    rows = list(csv.reader(open(os.path.join(src, f))))
    if len(rows) < 3:
        continue
    head, units, vals = rows[0], rows[1], rows[2]
    out = [f"# {f}: ncu --set full --clock-control none, one launch (kernel replay); selected metrics of the raw page"]
    for h, u, v in zip(head, units, vals):
        if h in ("Kernel Name", "Block Size", "Grid Size") or any(h.startswith(k) for k in KEEP):
            out.append(f"{h:90s} {v} {u}")
    open(f"{dst}_{f[:-8]}_ncu_full.txt", "w").write("\n".join(out) + "\n")
