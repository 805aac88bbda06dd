"""Shared helpers for the parity tests."""
import numpy as np

from mistra_b200.mechgen import mech as mechmod

# Comparison floor for diverged step sequences (SURVEY.md 8a trap 12):
# 1e-3 x (1 molecule cm^-3 = 1.66e-18 mol m^-3)
ATOL_FLOOR = 1.66e-21
RTOL = 1.0e-3


def random_cells(mech_name, ncell, seed):
    """Well-conditioned random cells for structure tests (not realistic chemistry):
    concentrations and rate constants log-uniform over a few decades, scaled so
    that the 10 s integration is stiff but finite."""
    m = mechmod.load(mech_name)
    r = np.random.default_rng(seed)
    var = 10.0 ** r.uniform(-12, -8, (ncell, m.nvar))
    fix = 10.0 ** r.uniform(-3, 1, (ncell, m.nfix))
    rc = np.empty((ncell, m.nreact))
    for i, facs in enumerate(m.reactions):
        order = sum(1 for f in facs if f[0] == "V")
        if order == 0:                       # pure source terms
            rc[:, i] = 10.0 ** r.uniform(-16, -13, ncell)
        elif order == 1:                     # first-order conversions: stiff (up to 1e2 1/s) but stable
            rc[:, i] = 10.0 ** r.uniform(-6, 2, ncell)
        else:                                # higher order: pseudo-first-order rate <= 1e-2 1/s at 1e-10
            rc[:, i] = 10.0 ** r.uniform(-6, -2, ncell) / (1e-10 ** (order - 1))
        nf = sum(1 for f in facs if f[0] == "F")
        rc[:, i] /= 10.0 ** nf               # fixed species are O(1e-3..1e1)
    return var, fix, rc


def rel_err(a, b, floor=ATOL_FLOOR):
    return np.abs(a - b) / (np.maximum(np.abs(a), np.abs(b)) + floor)
