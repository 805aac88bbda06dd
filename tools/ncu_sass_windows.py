#!/usr/bin/env python
"""Windowed view of `ncu --page source --csv --print-source sass`: samples, executed instructions, opcode mix
and stall mix per window of W SASS instructions (straight-line kernels: windows = phases).
usage: python tools/ncu_sass_windows.py sass.csv [W]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 400
hdr = rows[1]
ci = {}
for k, h in enumerate(hdr):
    ci.setdefault(h, k)
body = [r for r in rows[2:] if len(r) > ci["# Samples"]]
tot_s = sum(int(r[ci["# Samples"]] or 0) for r in body)
tot_i = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
print("SASS instructions %d, samples %d, warp instructions executed %d" % (len(body), tot_s, tot_i))


def opname(src):
    t = src.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else ""


for a in range(0, len(body), W):
    b = min(a + W, len(body))
    s = sum(int(r[ci["# Samples"]] or 0) for r in body[a:b])
    i = sum(int(r[ci["Instructions Executed"]] or 0) for r in body[a:b])
    if i == 0:
        continue
    ops = {}
    for r in body[a:b]:
        o = opname(r[ci["Source"]])
        ops[o] = ops.get(o, 0) + 1
    top = sorted(ops.items(), key=lambda t: -t[1])[:6]
    st = {}
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            st[h[6:]] = sum(int(r[ci[h]] or 0) for r in body[a:b])
    stt = sorted(st.items(), key=lambda t: -t[1])[:3]
    print("@%5d samples %5.2f%% inst %5.2f%% exec/instr %8.0f | %s | %s" % (
        a, 100 * s / tot_s, 100 * i / tot_i, i / (b - a), " ".join("%s:%d" % o for o in top),
        " ".join("%s:%.0f%%" % (k, 100 * v / max(1, s)) for k, v in stt)))
