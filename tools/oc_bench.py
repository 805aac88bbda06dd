"""A/B timing of the Ros3 kernels on the synthetic ensembles (tool, not product code).

  python tools/oc_bench.py [gas|aer|tot] [ncol] [spinup] [check]

MISTRA_KPP_ONCHIP=0 selects the default cell-per-thread kernel, 1 (default of this tool) the on-chip variant.  Prints cells/s, Ros3 steps/s,
the FP64 roofline fraction (SURVEY 8d flop table) and, with check > 0, the parity of `check`
cells of the timed output against the CPU oracle.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic  # noqa: E402

FLOPS = {0: (5443, 19263), 1: (14708, 200890), 2: (23950, 549260)}   # per accepted step, per attempt


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "aer"
    ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    spin = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    check = int(sys.argv[4]) if len(sys.argv) > 4 else 500
    cls = {"gas": synthetic.GasEnsemble, "aer": synthetic.AerEnsemble, "tot": synthetic.TotEnsemble}[name]
    ens = cls(ncol)
    mech = {"gas": 0, "aer": 1, "tot": 2}[name]
    variant = int(os.environ.get("MISTRA_KPP_ONCHIP", "1")) if mech < 2 else 0
    kpp.set_kernel(mech, variant)
    var = ens.var.copy()
    t0 = time.time()
    for _ in range(spin):
        rc = ens.rconst(var)
        var, ierr, stats, _, _ = kpp.integrate(mech, rc, ens.fix, var)
    rc = ens.rconst(var)
    print("%s: %d cells, spin-up %d steps %.1fs, ierr ok %.4f" % (name, ens.ncell, spin, time.time() - t0, (ierr == 1).mean()), flush=True)
    d_rc = torch.from_numpy(rc).cuda()
    d_fix = torch.from_numpy(np.ascontiguousarray(ens.fix)).cuda()
    d_var0 = torch.from_numpy(var).cuda()
    d_var = d_var0.clone()
    d_ierr = torch.zeros(ens.ncell, dtype=torch.int32, device="cuda")
    d_stats = torch.zeros((ens.ncell, 8), dtype=torch.int32, device="cuda")
    times = []
    ph = None
    if os.environ.get("OC_PHASES"):
        import ctypes as C
        ph = torch.zeros(16, dtype=torch.int64, device="cuda")
        L = kpp.library()
        L.mistra_kpp_oc_debug.argtypes = [C.c_void_p, C.c_longlong]
        L.mistra_kpp_oc_debug.restype = None
        L.mistra_kpp_oc_debug(C.c_void_p(ph.data_ptr()), 0)
    for it in range(4):
        d_var.copy_(d_var0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kpp.integrate_device(mech, d_rc, d_fix, d_var, ierr=d_ierr, stats=d_stats)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    st = d_stats.cpu().numpy()
    nstp, nacc, nrej = st[:, 2].sum(), st[:, 3].sum(), st[:, 4].sum()
    fa, ft = FLOPS[mech]
    flop = float(nacc) * fa + float(nstp) * ft
    peak = kpp.fp64_peak_tflops()
    print("%s onchip=%s: %.3f ms (%s), %.3f M cells/s, %.2f M steps/s, nstp/cell %.2f, nrej %d, %.2f TFLOP/s = %.3f of FP64 peak %.1f"
          % (name, os.environ.get("MISTRA_KPP_ONCHIP", "1"), ms, ",".join("%.1f" % t for t in times), ens.ncell / ms / 1e3,
             nstp / ms / 1e3, nstp / ens.ncell, nrej, flop / ms / 1e9, flop / ms / 1e9 / peak, peak), flush=True)
    if ph is not None:
        v = ph.cpu().numpy().astype(np.float64) / 4.0 / max(1, nstp)       # 4 timed launches
        names = ["jacprep", "fun K1", "hops", "ht", "tail_lu", "fwd frames", "tail fwd", "tail bwd", "bwd frames", "fun stage2 + rhs",
                 "combos/errnorm/ctl", "cell load/store"]
        tot = v[:12].sum()
        print("cycles per Ros3 step and cell (thread 0 of each block): total %.0f" % tot)
        for n_, x in zip(names, v):
            print("  %-20s %9.0f  %5.1f %%" % (n_, x, 100 * x / tot))
    if check:
        from oracle import kpp_oracle as ko
        idx = np.linspace(0, ens.ncell - 1, min(check, ens.ncell)).astype(np.int64)
        ref, ierr_o, stats_o, _, _ = ko.integrate(mech, rc[idx], ens.fix[idx], var[idx], nthreads=8)
        out = d_var.cpu().numpy()[idx]
        rel = np.abs(out - ref) / (np.abs(ref) + 1.66e-21)
        same = (st[idx, 2:5] == stats_o[:, 2:5]).all(axis=1)
        print("parity vs oracle on %d cells: max rel %.3e, max rel (same sequence) %.3e, same sequence %.3f, ierr equal %s"
              % (len(idx), rel.max(), rel[same].max() if same.any() else -1, same.mean(), (d_ierr.cpu().numpy()[idx] == ierr_o).all()), flush=True)


if __name__ == "__main__":
    main()
