"""Oracle against the committed golden vectors (CPU)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["gas_cells", "aer_cells", "tot_cells"])
def test_oracle_reproduces_golden(oracle, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out, ierr, stats, hexit, _ = oracle.integrate(int(g["mech"]), g["rconst"], g["fix"], g["var"])
    assert np.array_equal(ierr, g["ierr"]) and np.array_equal(stats, g["stats"])
    # same compiler, same flags -> bit-identical; allow a few ulp for other hosts' libm pow()
    assert np.allclose(out, g["var_out"], rtol=1e-12, atol=0.0)
    assert np.allclose(hexit, g["hexit"], rtol=1e-12)
