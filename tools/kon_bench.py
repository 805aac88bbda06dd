#!/usr/bin/env python
"""Small stand-alone driver of the condensation kernel (for timing / ncu captures):
python tools/kon_bench.py [layers] [iterations]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kon  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
it = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
grid = kon.kon_grid()
d = kon.synthetic_layers(grid, n)
keys = ("ffk", "totr", "dfdt", "feualt", "pp", "to", "tn", "xm1o", "xm1n", "kr")
t0 = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev) for k in keys}
st = torch.zeros(n, dtype=torch.int32, device=dev)
for i in range(it):
    t = {k: v.clone() for k, v in t0.items()}
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    kon.subkon_device(grid, 10.0, *[t[k] for k in keys], status=st)
    e[1].record()
    torch.cuda.synchronize()
    ms = e[0].elapsed_time(e[1])
    print("iter %d: subkon %.3f ms for %d layers = %.0f layers/s; mean iterations %.2f" % (
        i, ms, n, n / (ms * 1e-3), st.float().mean().item()))
if os.environ.get("MISTRA_KPP_LIB", "").endswith("_konprof.so"):
    import ctypes as C
    from mistra_b200 import kpp
    buf = (C.c_ulonglong * 8)()
    kpp.library().mistra_kon_prof(buf)
    v = list(buf)[:6]
    names = ["layer scalars + tile copy issue", "coefficient setup", "part A (flux)", "part B (ordered adds)", "secant update", "write back"]
    for nm, c in zip(names, v):
        print("%-34s %6.2f%%  %.0f cycles/layer" % (nm, 100.0 * c / sum(v), c / float(n * it)))
