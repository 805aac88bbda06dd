"""Mint the golden vectors of tests/golden/*.npz.

The reference has no golden vectors for this path and cannot be run here
(Fortran only), so these fixtures are outputs of the CPU oracle (the
statement-by-statement restatement of the reference path) on seeded inputs;
they pin the oracle and the CUDA path against regressions and travel to the GPU
box.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mistra_b200 import synthetic  # noqa: E402
from oracle import kpp_oracle as ko  # noqa: E402
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, mech, var, fix, rc):
    out, ierr, stats, hexit, texit = ko.integrate(mech, rc, fix, var)
    np.savez_compressed(os.path.join(HERE, name), mech=mech, var=var, fix=fix, rconst=rc, var_out=out,
                        ierr=ierr, stats=stats, hexit=hexit)
    print(name, var.shape, "ierr", np.unique(ierr), "nstp", stats[:, 2].min(), stats[:, 2].max())


def main():
    # gas: 24 cells of the synthetic ensemble (3 columns: layers spread over the column),
    # half of them after a 3-step spin-up (stiff quasi-steady radicals), half from cold start
    ens = synthetic.GasEnsemble(3, seed=synthetic.SEED)
    var = ens.var
    rc = ens.rconst(var)
    warm = var
    for _ in range(3):
        warm = ko.integrate(0, ens.rconst(warm), ens.fix, warm, nthreads=4)[0]
    pick = np.array([c * 148 + k for c in range(3) for k in (0, 40, 99, 147)])
    v = np.concatenate([var[pick], warm[pick]])
    f = np.concatenate([ens.fix[pick], ens.fix[pick]])
    r = np.concatenate([rc[pick], ens.rconst(warm)[pick]])
    save("gas_cells.npz", 0, v, f, r)
    # aer / tot: seeded well-conditioned random cells (structure test inputs)
    var, fix, rc = util.random_cells("aer", 8, 20261018)
    save("aer_cells.npz", 1, var, fix, rc)
    var, fix, rc = util.random_cells("tot", 4, 20261019)
    save("tot_cells.npz", 2, var, fix, rc)


if __name__ == "__main__":
    main()
