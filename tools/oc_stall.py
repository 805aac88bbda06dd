"""Hunt for the ~23 ms stalls of the on-chip gas kernel at some batch sizes: time per size, and per-phase cycle sums
(OC_PHASE_TIMERS build) for a slow and a fast size."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic
ens = synthetic.GasEnsemble(20)
var = ens.var
kpp.set_kernel(0, 0)
for _ in range(8):
    var = np.maximum(kpp.integrate(0, ens.rconst(var), ens.fix, var)[0], 0.0)
rc = ens.rconst(var)
d_rc, d_fix, d_var0 = torch.from_numpy(rc).cuda(), torch.from_numpy(np.ascontiguousarray(ens.fix)).cuda(), torch.from_numpy(var).cuda()
kpp.set_kernel(0, 1)
ph = torch.zeros(16 + 4 * 148, dtype=torch.int64, device="cuda")
L = kpp.library()
L.mistra_kpp_oc_debug.argtypes = [C.c_void_p, C.c_longlong]
L.mistra_kpp_oc_debug(C.c_void_p(ph.data_ptr()), 0)
def run(n, reps=3):
    work = d_var0[:n].clone()
    ts = []
    for it in range(reps):
        work.copy_(d_var0[:n]); ph.zero_(); ph[12] = 2**62; torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); kpp.integrate_device(0, d_rc[:n], d_fix[:n], work); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts, ph.cpu().numpy().copy()
slow = []
for n in [8, 300, 350, 512]:
    ts, p = run(n)
    flag = "SLOW" if min(ts) > 5 else ""
    print("n=%5d  ms %s %s" % (n, " ".join("%.2f" % t for t in ts), flag), flush=True)
    if flag or n in (8, 512):
        print("        phase cycles (block 0 thread 0 sums over all blocks):", p[:12].tolist(), flush=True)
        blk = p[16:].reshape(148, 4)
        nb = int((blk[:, 0] > 0).sum())
        dur = (blk[:nb, 1] - blk[:nb, 0]) * 1e-6
        print("        per block ms (start->end):", " ".join("%.2f" % d for d in dur[:40]), "| SM ids", blk[:nb, 2][:40].tolist(), flush=True)
        print("        global timer: first block start -> last block end %.3f ms; first start -> last start %.3f ms" % ((p[13] - p[12]) * 1e-6, (p[14] - p[12]) * 1e-6), flush=True)
