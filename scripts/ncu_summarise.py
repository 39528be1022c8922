"""Turns the captures of scripts/ncu_capture.sh into the text summaries committed under profiles/."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def summarise(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = OrderedDict()
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out[k] = f"{vals[i]} {units[i]}"
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    return name, out


def launch_shares(csv_path):
    agg = OrderedDict()
    for r in csv.DictReader(l for l in open(csv_path) if not l.startswith("==")):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit.startswith("ns") else v if unit.startswith("us") else v * 1000.0
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    lines = []
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n:5d} launches {us:10.1f} us {100 * us / tot:5.1f}%  avg {us / n:8.1f} us  {k}")
    return lines


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print("\n".join(launch_shares(sys.argv[2])))
    else:
        name, m = summarise(sys.argv[1])
        print("kernel:", name)
        for k, v in m.items():
            print(f"{k:75s} {v}")
