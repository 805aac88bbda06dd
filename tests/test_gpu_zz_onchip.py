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


def run_dev(kpp, cuda_device, rc, fix, var, strict=False, diag=True):
    """mistra_kpp_integrate_device on copies of the host arrays; returns numpy (var, ierr, stats, hexit, texit)."""
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    n = var.shape[0]
    vd = t(var)
    ie = torch.zeros(n, dtype=torch.int32, device=cuda_device) if diag else None
    sd = torch.zeros((n, 8), dtype=torch.int32, device=cuda_device) if diag else None
    hx = torch.zeros(n, dtype=torch.float64, device=cuda_device) if diag else None
    tx = torch.zeros(n, dtype=torch.float64, device=cuda_device) if diag else None
    kpp.integrate_device(1, t(rc), t(fix), vd, ierr=ie, stats=sd, hexit=hx, texit=tx, strict=strict)
    torch.cuda.synchronize()
    c = lambda x: None if x is None else x.cpu().numpy()
    return c(vd), c(ie), c(sd), c(hx), c(tx)


def test_handoff_of_long_cells(kpp, cuda_device, oracle):
    """mistra_kpp_set_handoff (device entry): cells that exceed the step budget in the cell-per-thread kernel are
    continued by the on-chip kernel from (T, H, counters).  Cold aer cells (~180 steps each): every cell is handed over
    after 12 attempts; the step history is that of an uninterrupted integration - identical in the strict build, within
    rounding in the product build - and cells that finish within the budget are not touched by the second pass."""
    ens = synthetic.AerEnsemble(2)
    rc = ens.rconst(ens.var)
    ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(1, rc, ens.fix, ens.var, nthreads=8)
    run = lambda *a, **k: run_dev(kpp, cuda_device, *a, **k)
    try:
        for strict in (False, True):
            kpp.set_kernel(1, 0, strict=strict)
            kpp.set_handoff(1, 0, strict=strict)
        plain, ierr_p, stats_p, hexit_p, texit_p = run(rc, ens.fix, ens.var)
        n0, n1 = kpp.launch_count_variant(0), kpp.launch_count_variant(1)
        kpp.set_handoff(1, 12)
        out, ierr, stats, hexit, texit = run(rc, ens.fix, ens.var)
        assert kpp.launch_count_variant(0) == n0 + 1 and kpp.launch_count_variant(1) == n1 + 1
        assert (stats_p[:, 2] > 12).all()                            # every cell exceeded the budget ...
        assert np.array_equal(ierr, ierr_p) and (ierr == 1).all() and np.array_equal(texit, texit_p)
        same = (stats == stats_p).all(axis=1)
        assert same.mean() >= 0.95                                   # ... and went on as if nothing had happened
        assert util.rel_err(out[same], plain[same]).max() <= 1e-5
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o, locked_tol=1e-5, hexit_rtol=1e-4)
        # the host-buffer entries hand over too (one pass after the last chunk, final rows put in place on the host)
        h_out, h_ierr, h_stats, h_hexit, h_texit = kpp.integrate(1, rc, ens.fix, ens.var)
        assert np.array_equal(h_out, out) and np.array_equal(h_stats, stats) and np.array_equal(h_ierr, ierr)
        assert np.array_equal(h_hexit, hexit) and np.array_equal(h_texit, texit)
        r1 = kpp.integrate_rates(1, ens.compact_rates(), ens.fix, ens.var)
        kpp.set_handoff(1, 0)
        r0 = kpp.integrate_rates(1, ens.compact_rates(), ens.fix, ens.var)
        kpp.set_handoff(1, 12)
        assert np.array_equal(r1[1], r0[1]) and (r1[2] == r0[2]).all(axis=1).mean() >= 0.95
        assert util.rel_err(r1[0], r0[0]).max() <= 1e-4 and util.rel_err(r1[0], ref).max() <= util.RTOL
        # strict build: both kernels reproduce the reference order, so the hand-off must be invisible
        kpp.set_handoff(1, 12, strict=True)
        outs, ierrs, statss, hexits, _ = run(rc, ens.fix, ens.var, strict=True)
        kpp.set_handoff(1, 0, strict=True)
        outp, ierrp, statsp, hexitp, _ = run(rc, ens.fix, ens.var, strict=True)
        assert np.array_equal(statss, stats_o) and np.array_equal(ierrs, ierr_o) and np.array_equal(statss, statsp)
        assert np.array_equal(hexits, hexitp) and util.rel_err(outs, outp).max() <= 1e-12
        # budgets around the step counts of a spun-up ensemble: nobody / some / everybody handed over
        var = ref
        for _ in range(2):
            var = kpp.integrate(1, ens.rconst(var), ens.fix, var)[0]
        rc2 = ens.rconst(var)
        base, ierr_b, stats_b, _, _ = run(rc2, ens.fix, var)
        for budget in (64, int(np.median(stats_b[:, 2])), 2):
            kpp.set_handoff(1, budget)
            o2, ierr2, stats2, _, _ = run(rc2, ens.fix, var)
            kpp.set_handoff(1, 0)
            assert np.array_equal(ierr2, ierr_b)
            moved = stats_b[:, 2] > budget                           # finished within the budget: never handed over
            assert np.array_equal(o2[~moved], base[~moved]) and np.array_equal(stats2[~moved], stats_b[~moved])
            assert util.rel_err(o2, base).max() <= 1e-4             # three steps after a cold start: rounding differences
                                                                     # between the kernels are still amplified (RTOL = 1e-3)
        kpp.set_handoff(1, 3)
        o3 = run(rc2, ens.fix, var, diag=False)[0]                   # the caller keeps no ierr / stats / hexit arrays
        assert util.rel_err(o3, base).max() <= 1e-4
        # failing cells are handed over like any other and fail the same way
        bad = var.copy()
        bad[5, 10] = np.nan
        kpp.set_handoff(1, 0)
        ierr_x = run(rc2, ens.fix, bad)[1]
        kpp.set_handoff(1, 4)
        ierr_y = run(rc2, ens.fix, bad)[1]
        assert np.array_equal(ierr_x, ierr_y) and ierr_y[5] < 0
        # a long batch (chunks on two streams, one hand-off pass at the end) with a few cold cells among spun-up ones
        reps = 1600                                                  # 313 600 cells: above the chunking threshold
        big_var, big_fix, big_rc = np.tile(var, (reps, 1)), np.tile(ens.fix, (reps, 1)), np.tile(rc2, (reps, 1))
        cold = np.arange(7, big_var.shape[0], 9973)
        big_var[cold], big_rc[cold] = ens.var[cold % ens.ncell], rc[cold % ens.ncell]
        kpp.set_handoff(1, 0)
        b0, ierr0, st0, _, _ = run(big_rc, big_fix, big_var)
        kpp.set_handoff(1, 12)
        b1, ierr1, st1, _, _ = run(big_rc, big_fix, big_var)
        assert np.array_equal(ierr0, ierr1) and (st0[cold, 2] > 12).all()
        quick = st0[:, 2] <= 12
        assert quick.mean() > 0.99 and np.array_equal(b1[quick], b0[quick]) and np.array_equal(st1[quick], st0[quick])
        assert (st1[~quick] == st0[~quick]).all(axis=1).mean() >= 0.9 and util.rel_err(b1, b0).max() <= 1e-4
        with pytest.raises(kpp.KppError):
            kpp.set_handoff(2, 5)                                    # tot has no on-chip kernel
    finally:
        for strict in (False, True):
            kpp.set_handoff(1, -1, strict=strict)
            kpp.set_kernel(1, 0, strict=strict)
