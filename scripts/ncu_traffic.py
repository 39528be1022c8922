#!/usr/bin/env python
"""Measures DRAM traffic per decoder-step launch for every BASELINE operating point (run on the GPU box):

    python scripts/ncu_traffic.py            # writes profiles/r2_decoder_traffic.json + the raw CSVs next to it

For each (model, batch): ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over three
mid-decode launches of decoder_step3_kernel (kernel replay; clocks untouched).  bench.py reads the JSON for
`roofline.traffic` instead of a constant.  Numbers under ncu are traffic evidence only, never bench values."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = [("tiny", 32), ("tiny", 1), ("base", 256), ("base_streaming", 64)]
KERNEL = os.environ.get("NCU_KERNEL", "decoder_step[34]_kernel")


def main():
    out = {}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for model, B in CONFIGS:
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
               "-k", f"regex:{KERNEL}", "-s", "30", "-c", "3", "--csv", sys.executable, os.path.join(ROOT, "scripts", "prof_step.py"), model, str(B)]
        env = dict(os.environ, MOONSHINE_B200_V4_COOP="0")  # ncu cannot replay a clustered + cooperative launch
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        text = r.stdout
        start = text.find('"ID"')
        if start < 0:
            print("ncu gave no table for", model, B, r.stderr[-400:], file=sys.stderr)
            continue
        with open(os.path.join(ROOT, "profiles", f"r2_decoder_traffic_{model}_b{B}.csv"), "w") as f:
            f.write(text[start:])
        per = {}
        for row in csv.DictReader(io.StringIO(text[start:])):
            v = float(row["Metric Value"].replace(",", ""))
            unit = row["Metric Unit"].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
            per.setdefault(row["ID"], {})[row["Metric Name"]] = v * scale
        per = {k: v for k, v in per.items() if all(x == x for x in v.values())}  # drop launches ncu could not measure
        n = len(per)
        if n == 0:
            print("ncu measured no launch for", model, B, file=sys.stderr)
            continue
        rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values()) / n
        wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values()) / n
        ns = sum(p.get("gpu__time_duration.sum", 0) for p in per.values()) / n
        out[f"{model}_b{B}"] = rd + wr
        out[f"{model}_b{B}_detail"] = {"read": rd, "write": wr, "duration_ns_under_ncu": ns, "launches": n, "kernel": KERNEL}
        print(model, B, out[f"{model}_b{B}_detail"], flush=True)
    with open(os.path.join(ROOT, "profiles", "r2_decoder_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
