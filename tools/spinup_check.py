"""How many spin-up chemistry steps does the synthetic ensemble need before the timed step is representative?
Prints mean Ros3 steps per cell / rejected steps of the NEXT step after 0, 6, 12, 30, 60 spin-up steps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200 import kpp, synthetic
for name, cls, mech in (("gas", synthetic.GasEnsemble, 0), ("aer", synthetic.AerEnsemble, 1)):
    ens = cls(200)
    var = ens.var
    rc = ens.rconst(var)
    marks = {0, 6, 12, 30, 60}
    for s in range(61):
        if s and s % 6 == 0:
            rc = ens.rconst(var)
        out, ierr, stats, _, _ = kpp.integrate(mech, rc, ens.fix, var)
        if s in marks:
            print("%s after %2d spin-up steps: next step takes %.3f Ros3 steps per cell (max %d), %d rejected, failed %d"
                  % (name, s, stats[:, 2].mean(), stats[:, 2].max(), stats[:, 4].sum(), (ierr != 1).sum()), flush=True)
        var = np.maximum(out, 0.0)
