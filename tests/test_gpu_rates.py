"""mistra_kpp_integrate_rates (include/mistra_kpp_rates.h): the rate constants formed on the device from compact
inputs - only the exchanged species of the NSPEC-indexed arrays travel - against the RCONST path (host
Update_RCONST_x + mistra_kpp_integrate) and the CPU oracle."""
import numpy as np
import pytest

from mistra_b200 import synthetic
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cls,mech", [(synthetic.GasEnsemble, 0), (synthetic.AerEnsemble, 1), (synthetic.TotEnsemble, 2)])
def test_compact_rates_path_matches_the_rconst_path(kpp, oracle, cuda_device, cls, mech):
    ens = cls(2)
    var = ens.var
    for step in range(2):
        rc = ens.rconst(var)                                   # host Update_RCONST_x
        ref, ierr_r, stats_r, _, _ = kpp.integrate(mech, rc, ens.fix, var)
        cr = ens.compact_rates()
        out, ierr, stats, hexit, texit, moved = kpp.integrate_rates(mech, cr, ens.fix, var)
        assert np.array_equal(ierr, ierr_r)
        # device exp / pow / log10 differ from libm by a few ulp per rate constant
        same = (stats[:, 2:5] == stats_r[:, 2:5]).all(axis=1)
        assert same.mean() >= 0.95
        assert util.rel_err(out[same], ref[same]).max() <= 1e-5
        assert util.rel_err(out, ref).max() <= util.RTOL
        full = rc.nbytes + ens.fix.nbytes + var.nbytes
        assert moved < full                                   # fewer bytes than shipping RCONST
        if mech > 0:
            assert moved < 0.65 * full
        var = np.maximum(ref, 0.0)


def test_lists_carry_only_exchanged_species(kpp, cuda_device):
    ens = synthetic.AerEnsemble(1)
    cr = ens.compact_rates()
    nspec = ens.m.nvar + ens.m.nfix
    for name in ("yhenry", "yxkmt", "ykef", "ykeb"):
        assert 0 < len(cr.idx[name]) <= 80 < nspec
    assert len(cr.idx["yxkmtd"]) <= 8 and len(cr.idx["yxeq"]) <= 8


def test_bad_lists_and_ragged_batches(kpp, cuda_device):
    ens = synthetic.AerEnsemble(1)
    cr = ens.compact_rates()
    out, ierr, _, _, _, _ = kpp.integrate_rates(1, cr, ens.fix, ens.var)
    assert (ierr == 1).all()
    cr.idx["yxkmt"][0] = 100000                               # species index out of range: rejected before any copy
    with pytest.raises(kpp.KppError):
        kpp.integrate_rates(1, cr, ens.fix, ens.var)
    with pytest.raises(kpp.KppError):
        kpp.integrate_rates(1, ens.compact_rates(), ens.fix[:5], ens.var[:5])
