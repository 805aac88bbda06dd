import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic
from oracle import kpp_oracle as ko
ens = synthetic.GasEnsemble(8)
var = ens.var
rc = ens.rconst(var)
ref, ierr_o, stats_o, hexit_o, _ = ko.integrate(0, rc, ens.fix, var, nthreads=8)
for variant in (0, 1, 1):
    kpp.set_kernel(0, variant)
    out, ierr, stats, hexit, _ = kpp.integrate(0, rc, ens.fix, var)
    bad = np.nonzero(ierr != ierr_o)[0]
    print("variant", variant, "ierr mismatches", len(bad), bad[:10], "ierr", ierr[bad[:10]], "oracle", ierr_o[bad[:10]])
    for c in bad[:5]:
        print("  cell", c, "stats", stats[c], "oracle", stats_o[c], "hexit", hexit[c], hexit_o[c])
    seq = (stats[:, 2:5] != stats_o[:, 2:5]).any(axis=1)
    print("  step-sequence mismatches", seq.sum(), "max rel", (np.abs(out-ref)/(np.abs(ref)+1.66e-21))[ierr == ierr_o].max())
