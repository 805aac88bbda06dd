"""Occupancy sweep of the difc / difp solve kernel on the device: MISTRA_DIFC_CTAS_PER_SM limits the
resident CTAs per SM (the xf values that wait for the backward sweep should stay in the L2).
usage (GPU box): for r in 16 8 6 4 3 2; do MISTRA_DIFC_CTAS_PER_SM=$r python tools/difc_sweep.py; done"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import difc as dm  # noqa: E402

dev = torch.device("cuda:0")
ncol, n = 2000, 150
dc = dm.synthetic_columns(ncol, n, seed=1)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dd = {k: t(dc[k]) for k in ("atkh", "w", "am3", "detw", "deta")}
rows = ((93, 93), (24, 24), (484, 484), (220, 220))
fd = [(torch.rand((ncol, n, r), dtype=torch.float64, device=dev) * dd["am3"][:, :, None], p) for r, p in rows]
ffp = torch.rand((500, n, 4900), dtype=torch.float64, device=dev)
fsp = torch.zeros((500, n), dtype=torch.float64, device=dev)
rho = dd["am3"][:500] / 35.0


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


m1 = timeit(lambda: dm.difc_device(60.0, dd["atkh"], dd["w"], dd["am3"], dd["detw"], dd["deta"], fd))
m2 = timeit(lambda: dm.difp_device(60.0, dd["atkh"][:500], dd["w"][:500], rho, dd["detw"], dd["deta"], ffp, fsp))
b1 = ncol * 821 * (2 * (n - 2) + 1) * 8
b2 = 500 * 4900 * 2 * (n - 1) * 8
ffp.mul_((torch.rand_like(ffp) < 0.05).double())          # 5 % of the bins populated, as real spectra are
for a, _ in fd:
    a[:, :, ::3] = 0.0                                     # species absent from the column
m3 = timeit(lambda: dm.difc_device(60.0, dd["atkh"], dd["w"], dd["am3"], dd["detw"], dd["deta"], fd))
m4 = timeit(lambda: dm.difp_device(60.0, dd["atkh"][:500], dd["w"][:500], rho, dd["detw"], dd["deta"], ffp, fsp))
print("sparse inputs: difc %.3f ms  difp %.3f ms" % (m3, m4))
print("CTAs/SM %s: difc %.3f ms (%.0f GB/s)  difp %.3f ms (%.0f GB/s)" % (
    os.environ.get("MISTRA_DIFC_CTAS_PER_SM", "16"), m1, b1 / m1 * 1e-6, m2, b2 / m2 * 1e-6))
