#!/usr/bin/env python
"""Small stand-alone driver of the 2-D bin redistribution kernels (for ncu captures):
python tools/bins_bench.py [layers] [iterations]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import bins  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
it = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
grid = bins.particle_grid()
d = bins.synthetic_layers(grid, n)
t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in d.items()}
ff0, si0, sl0 = t["ff"].clone(), t["sion1_new"].clone(), t["sl1"].clone()
sap = torch.zeros((n, 4), dtype=torch.float64, device=dev)
smp = torch.zeros_like(sap)
so = torch.zeros((n, 4, 9), dtype=torch.float64, device=dev)
for i in range(it):
    t["ff"].copy_(ff0); t["sion1_new"].copy_(si0); t["sl1"].copy_(sl0)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    bins.snapshot_device(grid, t["ff"], t["cm"], t["sion1"], sap, smp, so)
    e[1].record()
    bins.redistribute_device(grid, t["ff"], t["cm"], t["cw"], sap, smp, so, t["sion1_new"], t["sl1"])
    e[2].record()
    torch.cuda.synchronize()
    print("iter %d: snapshot %.3f ms, redistribute %.3f ms (%d layers)" % (i, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), n))
