#!/usr/bin/env python
"""Top stalled instructions per kernel from `ncu --page source --csv` output.
usage: python tools/ncu_source_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr) - 2:
                body.append(rows[j])
            j += 1
        ci = {h: k for k, h in enumerate(hdr)}
        s_all = ci["# Samples"]
        tot = sum(int(r[s_all] or 0) for r in body)
        print("=== %s\n    total samples %d, instructions %d" % (name[:100], tot, len(body)))
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stall_cols}
        print("    stall mix:", ", ".join("%s %.0f%%" % (h[6:], 100.0 * v / max(1, tot)) for h, v in sorted(agg.items(), key=lambda t: -t[1])[:6]))
        top = sorted(body, key=lambda r: -int(r[s_all] or 0))[:N]
        for r in top:
            st = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
            print("    %5.1f%%  %-70s %s" % (100.0 * int(r[s_all] or 0) / max(1, tot), r[ci["Source"]][:70],
                                            " ".join("%s:%d" % (b, a) for a, b in st)))
        i = j
    else:
        i += 1
