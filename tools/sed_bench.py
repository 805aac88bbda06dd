"""Device timing of the settling kernels (include/mistra_sed.h) on synthetic columns: CUDA events on the launching
stream, inputs resident, ff restored (untimed) before every pass.  python tools/sed_bench.py [ncol]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mistra_b200 import kon, sed  # noqa: E402

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nf, n, dt = 100, 150, 10.0
g = kon.kon_grid()
d = sed.synthetic_columns(g, ncol, seed=3)
t = lambda a, ty=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=ty)).cuda()
gd = dict(nka=g["nka"], nkt=g["nkt"], rq=t(g["rq"]), e=t(g["e"]), kw=t(g["kw"], np.int32))
dv = {k: t(v) for k, v in d.items()}
keep = {k: dv[k].clone() for k in ("ff", "sl1", "sion1", "diag")}
settled = ((d["ff"][:, 1:nf] * d["detw"][None, 1:nf, None, None]).sum(axis=1) > 1e-6).mean()


def timed(fn, reps=5):
    best = []
    for _ in range(reps):
        for k, v in keep.items():
            dv[k].copy_(v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    return float(np.median(best[1:]))


ms = timed(lambda: sed.sedp_device(gd, dt, nf, dv["detw"], dv["deta"], dv["t"], dv["p"], dv["vd"], dv["ff"], dv["diag"]))
rd = ncol * (nf - 1) * 4900 * 8
wr = ncol * (nf - 1) * 4900 * 8 * settled
print("sedp: %d columns (%.1f %% of the classes settle) %.3f ms = %.0f columns/s; algorithmic bytes %.2f GB -> %.0f GB/s"
      % (ncol, 100 * settled, ms, ncol / ms * 1e3, (rd + wr) / 1e9, (rd + wr) / ms / 1e6))
ms = timed(lambda: sed.sedl_device(dt, nf, 4, dv["detw"], dv["deta"], dv["t"], dv["p"], dv["rc"], dv["vt"], dv["vdm"],
                                   dv["sl1"], dv["sion1"]))
by = ncol * (nf - 1) * 4 * (121 + 55) * 8 * 2
print("sedl: %d columns %.3f ms = %.0f columns/s; algorithmic bytes %.3f GB -> %.0f GB/s" % (ncol, ms, ncol / ms * 1e3, by / 1e9, by / ms / 1e6))
