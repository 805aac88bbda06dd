#!/usr/bin/env python
"""Per-phase cycle breakdown of bins_redistribute_kernel from an instrumented build
(libmistra_kpp_prof.so, not part of the product)."""
import ctypes as C
import os
import subprocess
import sys

os.environ["MISTRA_KPP_LIB"] = "libmistra_kpp_prof.so"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "29600", "3"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bins_bench.py")).read())
from mistra_b200 import kpp  # noqa: E402
L = kpp.library()
buf = (C.c_ulonglong * 8)()
L.mistra_bins_prof(buf)
v = list(buf)
tot = sum(v)
names = ["phase0 den (+tile copy issue)", "phase1 ix/c0", "wait tile", "phase2 walk", "phase3 vc reduce", "tile store", "phase4 exchange"]
for n, c in zip(names, v):
    print("%-32s %6.2f%%  %.0f cycles/layer" % (n, 100.0 * c / tot, c / (29600.0 * 3)))
