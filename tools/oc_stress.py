"""Determinism stress of the on-chip kernels: the same batch many times, results must be bit-identical."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic
from tests import util
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cases = []
ens = synthetic.GasEnsemble(8); cases.append(("gas cold", 0, ens.rconst(ens.var), np.ascontiguousarray(ens.fix), ens.var.copy()))
ens = synthetic.AerEnsemble(2); cases.append(("aer cold", 1, ens.rconst(ens.var), np.ascontiguousarray(ens.fix), ens.var.copy()))
v, f, r = util.random_cells("aer", 1500, 11); cases.append(("aer random 1500", 1, r, f, v))
v, f, r = util.random_cells("gas", 5000, 12); cases.append(("gas random 5000", 0, r, f, v))
for strict in (False, True):
    for name, mech, rc, fix, var in cases:
        kpp.set_kernel(mech, 1, strict=strict)
        ref = None
        nbad = 0
        for it in range(N if not strict else max(3, N // 5)):
            out, ierr, stats, hexit, _ = kpp.integrate(mech, rc, fix, var, strict=strict)
            if ref is None:
                ref = (out.copy(), ierr.copy(), stats.copy())
                continue
            same = np.array_equal(out, ref[0], equal_nan=True) and np.array_equal(ierr, ref[1]) and np.array_equal(stats, ref[2])
            if not same:
                nbad += 1
                bad = np.nonzero((ierr != ref[1]) | (stats != ref[2]).any(axis=1) | ~((out == ref[0]) | (np.isnan(out) & np.isnan(ref[0]))).all(axis=1))[0]
                print("  %s strict=%s iteration %d: %d cells differ, first %s ierr %s vs %s stats %s vs %s" % (
                    name, strict, it, len(bad), bad[:8], ierr[bad[:4]], ref[1][bad[:4]], stats[bad[0]], ref[2][bad[0]]), flush=True)
        print("%s strict=%s: %d of %d repeats differ from the first run; ierr ok %.4f" % (name, strict, nbad, N, (ref[1] == 1).mean()), flush=True)
        kpp.set_kernel(mech, 0, strict=strict)
