"""The on-chip Ros3 kernel variant (csrc/ros3_onchip.inc: one persistent block per SM, cell slots in lockstep,
sparse head of the LU in shared memory, dense tail in registers) against the CPU oracle and against the default
cell-per-thread kernel, through the C ABI (mistra_kpp_set_kernel).  Same parity contract as test_gpu_parity.py;
the strict build must reproduce the oracle's accept/reject history in every cell."""
import os

import numpy as np
import pytest

from mistra_b200 import synthetic
from tests import util
from tests.test_gpu_parity import GOLD, compare   # noqa: E402  (module name: runs after the default-kernel suites)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def onchip(kpp, cuda_device):
    """Select the on-chip variant for gas and aer in both builds; restore the default afterwards."""
    for strict in (False, True):
        for mech in (0, 1):
            kpp.set_kernel(mech, 1, strict=strict)
            assert kpp.get_kernel(mech, strict=strict) == 1
    yield kpp
    for strict in (False, True):
        for mech in (0, 1):
            kpp.set_kernel(mech, 0, strict=strict)


def test_tot_has_no_onchip_kernel(kpp, cuda_device):
    with pytest.raises(kpp.KppError):
        kpp.set_kernel(2, 1)
    assert kpp.kernel_for(2, 1) == 0


def test_variant_by_batch_size(kpp, cuda_device):
    """Default: one cell (the box model) and small batches run on-chip, large batches on the cell-per-thread kernel."""
    for mech in (0, 1):
        kpp.set_kernel(mech, -1)
    try:
        assert kpp.get_kernel(1) == -1
        assert kpp.kernel_for(1, 1) == 1 and kpp.kernel_for(1, 10 ** 6) == 0
        assert kpp.kernel_for(0, 1) == 1 and kpp.kernel_for(0, 10 ** 6) == 0
        var, fix, rc = util.random_cells("aer", 40, 3)
        n1 = kpp.launch_count_variant(1)
        kpp.integrate(1, rc, fix, var)
        assert kpp.launch_count_variant(1) == n1 + 1
    finally:
        for mech in (0, 1):
            kpp.set_kernel(mech, 0)


@pytest.mark.parametrize("name", ["gas_cells", "aer_cells"])
def test_golden_vectors(onchip, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out, ierr, stats, hexit, _tx = onchip.integrate(int(g["mech"]), g["rconst"], g["fix"], g["var"])
    compare(out, g["var_out"], stats, g["stats"], ierr, g["ierr"], hexit, g["hexit"])


@pytest.mark.parametrize("name", ["gas_cells", "aer_cells"])
def test_strict_build_matches_golden_to_round_off(onchip, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out, ierr, stats, hexit, _tx = onchip.integrate(int(g["mech"]), g["rconst"], g["fix"], g["var"], strict=True)
    assert np.array_equal(ierr, g["ierr"])
    assert np.array_equal(stats, g["stats"])       # identical accept/reject history in every cell
    sig = np.abs(g["var_out"]) > 1e-30
    rel = (np.abs(out - g["var_out"]) / np.maximum(np.abs(g["var_out"]), 1e-300))[sig]
    assert rel.max() <= 1e-12, rel.max()


def test_gas_cold_start_with_rejections(onchip, oracle):
    ens = synthetic.GasEnsemble(8)                 # 1184 cells, cold start: many steps + rejections
    var = ens.var
    for step in range(2):
        rc = ens.rconst(var)
        ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(0, rc, ens.fix, var, nthreads=8)
        out, ierr, stats, hexit, _tx = onchip.integrate(0, rc, ens.fix, var)
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o)
        var = ref


def test_aer_cold_start(onchip, oracle):
    ens = synthetic.AerEnsemble(2)                 # 196 cells, ~180 steps in the first call
    var = ens.var
    rc = ens.rconst(var)
    ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(1, rc, ens.fix, var, nthreads=8)
    out, ierr, stats, hexit, _tx = onchip.integrate(1, rc, ens.fix, var)
    compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o, locked_tol=1e-5, hexit_rtol=1e-4)
    outs, ierrs, statss, _, _ = onchip.integrate(1, rc, ens.fix, var, strict=True)
    assert np.array_equal(statss, stats_o) and np.array_equal(ierrs, ierr_o)


def test_more_cells_than_slots_ragged_and_empty(onchip, oracle):
    """More cells than the 148 x 5 (aer) resident slots, a batch smaller than one block's slots, an empty batch."""
    var, fix, rc = util.random_cells("aer", 1500, 11)
    ref, ierr_o, stats_o, _, _ = oracle.integrate(1, rc, fix, var, nthreads=8)
    out, ierr, stats, _, _ = onchip.integrate(1, rc, fix, var)
    assert np.array_equal(ierr, ierr_o)
    assert util.rel_err(out, ref).max() <= util.RTOL
    out3, ierr3, _, _, _ = onchip.integrate(1, rc[:3], fix[:3], var[:3])
    assert np.array_equal(out3, out[:3]) and np.array_equal(ierr3, ierr[:3])
    out0, ierr0, _, _, _ = onchip.integrate(1, rc[:0], fix[:0], var[:0])
    assert out0.shape == (0, var.shape[1]) and ierr0.shape == (0,)


def test_failing_cells_and_step_limit(onchip, oracle):
    """NaN / Inf cells end with ierr = -7 (SURVEY 8a trap 9), a step limit of 3 with ierr = -6 and the partially
    advanced VAR; the healthy cells of the same batch are unaffected (every slot goes through every attempt of
    its block, whatever its cell does)."""
    var, fix, rc = util.random_cells("gas", 70, 6)
    var[3, 10] = np.nan
    var[40, 0] = np.inf
    ref, ierr_o, stats_o, _, _ = oracle.integrate(0, rc, fix, var)
    out, ierr, stats, _, _ = onchip.integrate(0, rc, fix, var)
    assert np.array_equal(ierr, ierr_o) and ierr[3] == -7 and ierr[40] == -7
    good = ierr == 1
    assert good.sum() == 68 and util.rel_err(out[good], ref[good]).max() <= util.RTOL
    o, oo = onchip.default_opts(max_steps=3), oracle.default_opts(max_steps=3)
    ref, ierr_o, stats_o, _, _ = oracle.integrate(0, rc[:8], fix[:8], var[8:16], opts=oo)
    out, ierr, stats, _, _ = onchip.integrate(0, rc[:8], fix[:8], var[8:16], opts=o)
    assert (ierr == -6).all() and np.array_equal(ierr, ierr_o) and np.array_equal(stats[:, 2], stats_o[:, 2])
    assert util.rel_err(out, ref).max() <= util.RTOL


def test_same_results_as_the_default_kernel(kpp, cuda_device):
    """Both variants on the same spun-up ensemble: identical step sequences, values within FMA-contraction noise."""
    ens = synthetic.AerEnsemble(4)
    var = ens.var
    for _ in range(2):
        var = kpp.integrate(1, ens.rconst(var), ens.fix, var)[0]
    rc = ens.rconst(var)
    kpp.set_kernel(1, 0)
    try:
        a, ierr_a, st_a, _, _ = kpp.integrate(1, rc, ens.fix, var)
        kpp.set_kernel(1, 1)
        b, ierr_b, st_b, _, _ = kpp.integrate(1, rc, ens.fix, var)
    finally:
        kpp.set_kernel(1, 0)
    assert np.array_equal(ierr_a, ierr_b)
    assert (st_a[:, 2:5] == st_b[:, 2:5]).all(axis=1).mean() >= 0.99
    assert util.rel_err(a, b).max() <= 1e-5
