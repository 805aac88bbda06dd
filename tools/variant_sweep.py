"""Throughput of the two Ros3 kernel variants against the batch size (device-resident batches, CUDA events):
where the on-chip kernel (few hundred cells in flight) hands over to the cell-per-thread kernel (needs tens of
thousands of cells).  Sets the crossovers KPP_ONCHIP_MAX_CELLS_* of csrc/kpp_api.cu."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic

for name, cls, mech, cols in (("gas", synthetic.GasEnsemble, 0, 1400), ("aer", synthetic.AerEnsemble, 1, 2100)):
    ens = cls(cols)
    var = ens.var
    kpp.set_kernel(mech, 0)
    for _ in range(12):
        var = np.maximum(kpp.integrate(mech, ens.rconst(var), ens.fix, var)[0], 0.0)
    rc = ens.rconst(var)
    d_rc, d_fix, d_var0 = torch.from_numpy(rc).cuda(), torch.from_numpy(np.ascontiguousarray(ens.fix)).cuda(), torch.from_numpy(var).cuda()
    perm = torch.randperm(ens.ncell, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    d_rc, d_fix, d_var0 = d_rc[perm].contiguous(), d_fix[perm].contiguous(), d_var0[perm].contiguous()   # mix the layers
    sizes = [1, 8, 64, 512, 740 if mech else 2220, 2048, 4096, 8192, 16384, 32768, 65536, 131072, ens.ncell]
    print("%s: cells, ms cell-per-thread, ms on-chip, ratio (on-chip faster if > 1)" % name, flush=True)
    for n in sizes:
        n = min(n, ens.ncell)
        ms = {}
        for variant in (0, 1):
            kpp.set_kernel(mech, variant)
            work = d_var0[:n].clone()
            ts = []
            for it in range(4):
                work.copy_(d_var0[:n])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                kpp.integrate_device(mech, d_rc[:n], d_fix[:n], work)
                e1.record()
                torch.cuda.synchronize()
                if it:
                    ts.append(e0.elapsed_time(e1))
            ms[variant] = float(np.median(ts))
        print("  %7d  %9.3f  %9.3f  %6.2f   (%.3f / %.3f M cells/s)" % (n, ms[0], ms[1], ms[0] / ms[1], n / ms[0] / 1e3, n / ms[1] / 1e3), flush=True)
    kpp.set_kernel(mech, -1)
