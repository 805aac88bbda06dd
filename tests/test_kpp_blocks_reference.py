"""The CPU oracle's Fun_x, Jac_SP_x and KppSolve_x against results obtained by executing the reference's own Fortran
statements (tests/golden/make_kpp_blocks_reference.py: reads SUBROUTINE Fun_x / Jac_SP_x / KppSolve_x from gas.f / aer.f /
tot.f, no code shared with mechgen/extract.py or oracle/emit_oracle.py).  Bit for bit: same statements, same order,
binary64 without fused multiply-add, default-REAL coefficients as binary32 values.  Together with
tests/test_rconst_reference.py this pins the oracle's generated blocks by an independent reading of the reference; the
hand-restated control flow of RosenbrockIntegrator_x stays pinned by structure only."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("mech,name", [(0, "gas"), (1, "aer"), (2, "tot")])
def test_oracle_blocks_equal_the_reference_statements(oracle, mech, name):
    g = np.load(os.path.join(GOLD, "kpp_blocks_reference_%s.npz" % name))
    for c in range(g["var"].shape[0]):
        V, F, R = g["var"][c], g["fix"][c], g["rconst"][c]
        assert np.array_equal(oracle.fun(mech, V, F, R, f32=1), g["vdot"][c])          # Fun_x
        assert np.array_equal(oracle.jac(mech, V, F, R, f32=1), g["jvs"][c])           # Jac_SP_x (fill-in slots = 0)
        assert np.array_equal(oracle.solve(mech, g["lu"][c], g["xin"][c]), g["xout"][c])   # KppSolve_x
