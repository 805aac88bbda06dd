"""Development aid: run ONE attempt of the on-chip Ros3 kernel (OC_DEBUG build) on the GPU, dump the block's
shared memory and tail registers at checkpoints and compare each with the CPU emulation of the same
instruction streams (mistra_b200/mechgen/onchip.py).

  python -m mistra_b200.build --variant ocdbg -DOC_DEBUG --units=kpp_onchip_g.cu,kpp_onchip_a.cu
  MISTRA_KPP_LIB=libmistra_kpp_ocdbg.so python tools/oc_debug.py gas|aer [cell]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic  # noqa: E402
from mistra_b200.mechgen import mech as mechmod, onchip  # noqa: E402

GAMMA1 = 0.43586652150845899941601945119356
C21 = -0.10156171083877702091975600115545e+01
C31 = 0.40759956452537699824805835358067e+01
C32 = 0.92076794298330791242156818474003e+01
NCHK = 13
TOL = 0.0 if os.environ.get("OC_STRICT") == "1" else 1e-9


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "gas"
    cell = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    mech = {"gas": 0, "aer": 1}[name]
    m = mechmod.load(name)
    strict = "strict" in os.environ.get("MISTRA_KPP_LIB", "") or os.environ.get("OC_STRICT") == "1"
    p = onchip.Plan(m, onchip.TAIL[name], strict=strict)
    ens = (synthetic.GasEnsemble if name == "gas" else synthetic.AerEnsemble)(1)
    var = ens.var.copy()
    rc = ens.rconst(var)
    fix = np.ascontiguousarray(ens.fix)
    per = p.smem_doubles + p.NT * p.R * 32
    if not torch.cuda.is_available():          # dry run of the emulation only
        D = np.full((NCHK, per), np.nan)
        return emulate(p, D, var, fix, rc, cell)
    buf = torch.full((NCHK, per), float("nan"), dtype=torch.float64, device="cuda")
    L = kpp.library()
    kpp.set_kernel(mech, 1)
    L.mistra_kpp_oc_debug.argtypes = [C.c_void_p, C.c_longlong]
    L.mistra_kpp_oc_debug.restype = None
    L.mistra_kpp_oc_debug(C.c_void_p(buf.data_ptr()), cell)
    d_rc, d_fix, d_var = torch.from_numpy(rc).cuda(), torch.from_numpy(fix).cuda(), torch.from_numpy(var).cuda()
    d_stats = torch.zeros((ens.ncell, 8), dtype=torch.int32, device="cuda")
    kpp.integrate_device(mech, d_rc, d_fix, d_var, stats=d_stats)
    torch.cuda.synchronize()
    D = buf.cpu().numpy()
    print("stats of the cell:", d_stats.cpu().numpy()[cell])
    emulate(p, D, var, fix, rc, cell)


def emulate(p, D, var, fix, rc, cell):
    # ---- emulation of the first attempt -------------------------------------------------------------
    H = 1e-3
    ghinv = 1.0 / (H * GAMMA1)
    S = p.new_smem()
    S[p.O_Y:p.O_Y + p.n] = var[cell]
    p.emu_set_consts(S, fix[cell], 1)
    regions = [("G", p.O_G, p.NG), ("ZERO", p.ZERO, 1), ("Y", p.O_Y, p.n), ("CY", p.O_CY, p.nc), ("T1", p.O_T1, p.n), ("CT", p.O_CT, p.nc),
               ("K1", p.O_K1, p.n), ("K2", p.O_K2, p.n), ("K3", p.O_K3, p.n), ("EX", p.O_EX, p.T)]

    def cmp(k, S, a, what, only=None):
        g = D[k]
        gs, ga = g[:p.smem_doubles], g[p.smem_doubles:].reshape(p.NT, p.R, 32)
        if np.isnan(gs).all():
            print("checkpoint %2d (%s): NOT WRITTEN" % (k, what))
            return
        msgs = []
        for nm, o, ln in regions:
            if only and nm not in only:
                continue
            e, r = S[o:o + ln], gs[o:o + ln]
            ok = np.isfinite(e)
            if not ok.any():
                continue
            with np.errstate(all="ignore"):
                rel = np.abs(e[ok] - r[ok]) / np.maximum(np.abs(e[ok]), 1e-300)
            rel = np.where(np.isfinite(rel), rel, np.inf)
            nbad = int((rel > TOL).sum())
            msgs.append("%s %s%d/%d max %.2e" % (nm, "BAD " if nbad else "", nbad, int(ok.sum()), rel.max() if rel.size else 0))
            if nbad and nbad <= 12:
                ii = np.nonzero(ok)[0][rel > TOL]
                msgs.append("[" + ", ".join("%d: emu %.6e gpu %.6e" % (i, e[i], r[i]) for i in ii[:12]) + "]")
            elif nbad:
                ii = np.nonzero(ok)[0][rel > TOL]
                msgs.append("[first bad idx %s ...]" % ii[:16].tolist())
        if a is not None and (only is None or "a" in only):
            with np.errstate(all="ignore"):
                rel = np.abs(a - ga) / np.maximum(np.abs(a), 1e-300)
            rel = np.where(np.isfinite(rel), rel, np.inf)
            nbad = int((rel > TOL).sum())
            msgs.append("a %s%d/%d max %.2e" % ("BAD " if nbad else "", nbad, a.size, rel.max()))
            if nbad:
                ii = np.argwhere(rel > TOL)[:10]
                msgs.append("[" + ", ".join("t%d q%d s%d: emu %.4e gpu %.4e" % (t, q, s, a[t, q, s], ga[t, q, s]) for t, q, s in ii) + "]")
        print("checkpoint %2d (%s): %s" % (k, what, " | ".join(msgs)))

    a, sing = p.emu_jacprep(S, rc[cell], ghinv)
    cmp(0, S, a, "after Jac_SP + matrix preparation", only=("G", "ZERO", "Y", "CY", "CT", "a"))
    p.emu_fun(S, p.O_Y, p.O_K1, rc[cell])
    cmp(1, S, a, "after Fun (K1)", only=("G", "K1", "Y", "a"))
    p.emu_hops(S)
    cmp(2, S, a, "after the head elimination", only=("G", "a"))
    p.emu_ht(S, a)
    cmp(3, S, a, "after the head pivots on the tail", only=("G", "a"))
    p.emu_tail_lu(a)
    cmp(4, S, a, "after the tail elimination", only=("G", "a"))
    T, h = p.T, p.h

    def A(r, c):
        return (32 * (c // 32) + (r % 32), r // 32, c % 32)

    def solve(xb, ks):
        xpb = p.O_EX
        S[xpb:xpb + T] = 0.0
        p.emu_frames(p.fwd_stream, p.fwd_nchunk, S, xb, xpb)
        if ks:
            cmp(ks[0], S, None, "stage 1 forward frames", only=("K1", "EX"))
        X = S[xb:xb + p.n]
        for r in range(T):
            if p.fwd_partial[r]:
                X[h + r] = X[h + r] + S[xpb + r]
        for c in range(T):
            for r in range(c + 1, T):
                X[h + r] = X[h + r] - a[A(r, c)] * X[h + c]
        if ks:
            cmp(ks[1], S, None, "stage 1 tail forward", only=("K1",))
        if p.strict:
            for r in range(T - 1, -1, -1):
                acc = X[h + r]
                for c in range(r + 1, T):
                    acc = acc - a[A(r, c)] * X[h + c]
                X[h + r] = acc / a[A(r, r)]
        else:
            for c in range(T - 1, -1, -1):
                X[h + c] = X[h + c] * a[A(c, c)]
                for r in range(c):
                    X[h + r] = X[h + r] - a[A(r, c)] * X[h + c]
        if ks:
            cmp(ks[2], S, None, "stage 1 tail backward", only=("K1",))
        p.emu_frames(p.bwd_stream, p.bwd_nchunk, S, xb, xpb)
        if ks:
            cmp(ks[3], S, None, "stage 1 backward frames", only=("K1",))

    solve(p.O_K1, (5, 6, 7, 8))
    n = p.n
    S[p.O_T1:p.O_T1 + n] = S[p.O_Y:p.O_Y + n] + S[p.O_K1:p.O_K1 + n]
    k1 = S[p.O_K1:p.O_K1 + n].copy()
    # Fun(T1) -> into a temporary (the kernel overwrites T1 with Fcn)
    tmp = p.new_smem()
    tmp[:] = S
    p.emu_fun(tmp, p.O_T1, p.O_T1, rc[cell])
    fcn = tmp[p.O_T1:p.O_T1 + n].copy()
    S[p.O_T1:p.O_T1 + n] = fcn
    S[p.O_K2:p.O_K2 + n] = fcn + (C21 / H) * k1
    S[p.O_K3:p.O_K3 + n] = fcn + (C31 / H) * k1
    cmp(9, S, None, "stage 2 right-hand sides", only=("T1", "K1", "K2", "K3"))
    solve(p.O_K2, None)
    cmp(10, S, None, "stage 2 solved", only=("K2",))
    S[p.O_K3:p.O_K3 + n] = S[p.O_K3:p.O_K3 + n] + (C32 / H) * S[p.O_K2:p.O_K2 + n]
    cmp(11, S, None, "stage 3 right-hand side", only=("K3",))
    solve(p.O_K3, None)
    cmp(12, S, None, "stage 3 solved", only=("K3",))


if __name__ == "__main__":
    main()
