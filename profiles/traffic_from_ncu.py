#!/usr/bin/env python
"""Turn an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum) of a bench.py run into
per-kernel averages: launches, mean DRAM bytes per launch, mean duration.
    python profiles/traffic_from_ncu.py gpurun_out/r02_launches.csv"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; ix = {h: i for i, h in enumerate(H)}
acc = defaultdict(lambda: defaultdict(list))
for r in rows[hdr + 1:]:
    if len(r) < len(H): continue
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
    acc[(r[ix["Kernel Name"]][:110], r[ix["Grid Size"]] if "Grid Size" in ix else "")][r[ix["Metric Name"]]].append(v * scale)
print(f"{'kernel':110s} {'launches':>8s} {'dram B/launch':>14s} {'ms/launch':>10s}")
for (k, g), m in sorted(acc.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    n = len(m.get("gpu__time_duration.sum", []))
    rd = sum(m.get("dram__bytes_read.sum", [0])) / max(n, 1); wr = sum(m.get("dram__bytes_write.sum", [0])) / max(n, 1)
    t = sum(m.get("gpu__time_duration.sum", [0])) / max(n, 1)
    print(f"{k:110s} {n:8d} {rd + wr:14.4e} {t * 1e3:10.4f}")
