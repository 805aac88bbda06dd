#!/usr/bin/env python
"""Per-phase SM-cycle breakdown of ros3_kernel_<x> from an instrumented experiment build
(python -m mistra_b200.build --variant pt -DKPP_PHASE_TIMERS --units=kpp_mech_a.cu,...);
not part of the product.  usage: tools/phase_times.py [mech=aer] [ncol=500]"""
import ctypes as C
import os
import sys
import time

os.environ.setdefault("MISTRA_KPP_LIB", "libmistra_kpp_pt.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mistra_b200 import kpp, synthetic  # noqa: E402

mech = sys.argv[1] if len(sys.argv) > 1 else "aer"
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 500
cls = {"gas": synthetic.GasEnsemble, "aer": synthetic.AerEnsemble, "tot": synthetic.TotEnsemble}[mech]
mi = {"gas": 0, "aer": 1, "tot": 2}[mech]
ens = cls(ncol)
dev = torch.device("cuda:0")
rc = torch.from_numpy(ens.rconst()).to(dev)
fix = torch.from_numpy(np.ascontiguousarray(ens.fix)).to(dev)
var0 = torch.from_numpy(ens.var).to(dev)
stats = torch.zeros((ens.ncell, 8), dtype=torch.int32, device=dev)
L = kpp.library()
fn = getattr(L, "mistra_kpp_phase_" + "gat"[mi])
buf = (C.c_ulonglong * 8)()
# spin up a few steps so that the timed call sees the quasi-steady state
var = var0.clone()
for _ in range(3):
    kpp.integrate_device(mi, rc, fix, var)
torch.cuda.synchronize()
fn(buf, 1)
v1 = var.clone()
t0 = time.time()
kpp.integrate_device(mi, rc, fix, v1, stats=stats)
torch.cuda.synchronize()
dt = time.time() - t0
fn(buf, 1)
v = list(buf)[:7]
tot = float(sum(v))
nstp = int(stats[:, 2].sum().item())
names = ["jacprep", "decomp", "fun<0>", "solve<1>", "fun<1>", "solve<2>", "solve<3>"]
print("%s: %d cells, %d Ros3 steps, %.1f ms -> %.3g cells/s" % (mech, ens.ncell, nstp, dt * 1e3, ens.ncell / dt))
for n, c in zip(names, v):
    print("%-10s %6.2f%%  %9.0f warp-cycles per 32 cell-steps" % (n, 100.0 * c / tot, c / (nstp / 32.0)))
