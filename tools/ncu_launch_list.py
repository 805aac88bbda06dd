#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list.
usage: python tools/ncu_launch_list.py x.csv > profiles/xxx.txt"""
import csv
import sys
from collections import defaultdict

tot, cnt = defaultdict(float), defaultdict(int)
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
    tot[r["Kernel Name"]] += v
    cnt[r["Kernel Name"]] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
    print("%6.2f%% %4d launches %10.3f ms  %s" % (100 * v / s, cnt[k], v, k[:110]))
