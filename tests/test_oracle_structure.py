"""The CPU oracle against size-independent properties (CPU only).

The reference ships no tests or golden vectors for this path, so
the oracle - a statement-by-statement restatement generated from the reference's
own KPP output - is pinned by structure: Jac_SP is the derivative of Fun, the
unrolled KppSolve inverts what KppDecomp factorises, the integrator converges
with the tolerance and agrees with an independent implicit solver.  (The
stoichiometric matrices have full row rank - lumped products and ignored species -
so there is no exact linear invariant to test.)"""
import numpy as np
import pytest

from mistra_b200.mechgen import mech as mechmod
from tests import util

MECHS = list(enumerate(mechmod.MECH_NAMES))


@pytest.mark.parametrize("mi,name", MECHS)
def test_query_and_names(oracle, mi, name):
    m = mechmod.load(name)
    assert oracle.query(mi) == (m.nvar, m.nfix, m.nreact, m.lu_nonzero)
    assert [oracle.spc_name(mi, i) for i in range(m.nspec)] == m.spc_names
    assert oracle.spc_name(mi, m.nspec) is None


@pytest.mark.parametrize("mi,name", MECHS)
def test_jac_is_derivative_of_fun(oracle, mi, name):
    m = mechmod.load(name)
    r = np.random.default_rng(7 + mi)
    V, F, R = r.uniform(0.5, 1.5, m.nvar), r.uniform(0.5, 1.5, m.nfix), r.uniform(0.5, 1.5, m.nreact)
    J = oracle.jac(mi, V, F, R)
    Jd = np.zeros((m.nvar, m.nvar))
    for j in range(m.nvar):
        h = 1e-4 * V[j]
        Vp, Vm = V.copy(), V.copy()
        Vp[j] += h
        Vm[j] -= h
        Jd[:, j] = (oracle.fun(mi, Vp, F, R) - oracle.fun(mi, Vm, F, R)) / (2 * h)
    Js = np.zeros_like(Jd)
    Js[m.row_of, m.icol] = J
    assert np.abs(Js - Jd).max() <= 1e-9 * np.abs(Jd).max()
    outside = np.ones_like(Jd, dtype=bool)
    outside[m.row_of, m.icol] = False
    assert np.abs(Jd[outside]).max() == 0.0


@pytest.mark.parametrize("mi,name", MECHS)
def test_decomp_solve_inverts_the_prepared_matrix(oracle, mi, name):
    m = mechmod.load(name)
    r = np.random.default_rng(11 + mi)
    V, F, R = r.uniform(0.5, 1.5, m.nvar), r.uniform(0.5, 1.5, m.nfix), r.uniform(0.5, 1.5, m.nreact)
    G = -oracle.jac(mi, V, F, R)
    G[m.diag[:m.nvar]] += 1.0 / (1e-2 * 0.43586652150845899941601945119356)
    LU, ier = oracle.decomp(mi, G)
    assert ier == 0
    b = r.normal(size=m.nvar)
    x = oracle.solve(mi, LU, b)
    Gd = np.zeros((m.nvar, m.nvar))
    Gd[m.row_of, m.icol] = G
    assert np.abs(Gd @ x - b).max() <= 1e-12 * max(1.0, np.abs(b).max())
    assert np.allclose(x, np.linalg.solve(Gd, b), rtol=1e-10, atol=1e-14)


def test_decomp_flags_zero_diagonal(oracle):
    m = mechmod.load("gas")
    G = np.ones(m.lu_nonzero)
    G[m.diag[4]] = 0.0
    _, ier = oracle.decomp(0, G)
    assert ier == 5    # 1-based row, tested before row 5 is eliminated (gas.f:6156)


@pytest.mark.parametrize("mi,name", MECHS)
def test_integrator_statistics_and_exit(oracle, mi, name):
    var, fix, rc = util.random_cells(name, 6, 100 + mi)
    out, ierr, stats, hexit, texit = oracle.integrate(mi, rc, fix, var)
    assert (ierr == 1).all() and np.allclose(texit, 10.0)
    nfun, njac, nstp, nacc, nrej, ndec, nsol, nsng = stats.T
    assert (nacc >= 7).all()                      # growth cap x6 from Hstart=1e-3
    assert (nfun == 2 * nacc + nstp).all() and (njac == nacc).all()
    assert (ndec == nstp).all() and (nsol == 3 * nstp).all() and (nsng == 0).all()
    assert (nstp >= nacc + nrej).all()
    assert np.isfinite(out).all() and (hexit > 0).all()


def test_zero_length_interval_returns_input(oracle):
    var, fix, rc = util.random_cells("gas", 2, 5)
    out, ierr, stats, hexit, texit = oracle.integrate(0, rc, fix, var, t0=3.0, t1=3.0)
    assert (ierr == 1).all() and (stats == 0).all() and np.array_equal(out, var)


def test_bad_tolerances_are_rejected(oracle):
    var, fix, rc = util.random_cells("gas", 1, 5)
    o = oracle.default_opts(rtol=2.0)
    _, ierr, *_ = oracle.integrate(0, rc, fix, var, opts=o)
    assert ierr[0] == -5
    o = oracle.default_opts(max_steps=3)
    _, ierr, stats, *_ = oracle.integrate(0, rc, fix, var, opts=o)
    assert ierr[0] == -6 and stats[0, 2] == 4     # Nstp > Max_no_steps test precedes the step


def test_nan_input_fails_with_step_too_small(oracle):
    var, fix, rc = util.random_cells("gas", 1, 5)
    var[0, 10] = np.nan
    _, ierr, *_ = oracle.integrate(0, rc, fix, var)
    assert ierr[0] == -7                          # SURVEY 8a trap 9


@pytest.mark.parametrize("mi,name", MECHS)
def test_self_convergence_with_tolerance(oracle, mi, name):
    var, fix, rc = util.random_cells(name, 3, 200 + mi)
    ref = oracle.integrate(mi, rc, fix, var, opts=oracle.default_opts(rtol=1e-9))[0]
    errs = []
    for rtol in (1e-3, 1e-5):
        out = oracle.integrate(mi, rc, fix, var, opts=oracle.default_opts(rtol=rtol))[0]
        sig = np.abs(ref) > 1e-16
        errs.append(np.abs(out - ref)[sig].max() / np.abs(ref)[sig].max())
    assert errs[1] < errs[0] or errs[0] < 1e-8
    assert errs[0] < 5e-2 and errs[1] < 5e-4


def test_agrees_with_scipy_radau_on_gas_cells(oracle):
    from scipy.integrate import solve_ivp
    m = mechmod.load("gas")
    var, fix, rc = util.random_cells("gas", 2, 300)
    out = oracle.integrate(0, rc, fix, var, opts=oracle.default_opts(rtol=1e-7))[0]
    for c in range(2):
        def f(t, y):
            return oracle.fun(0, y, fix[c], rc[c])

        def jac(t, y):
            J = np.zeros((m.nvar, m.nvar))
            J[m.row_of, m.icol] = oracle.jac(0, y, fix[c], rc[c])
            return J
        sol = solve_ivp(f, (0.0, 10.0), var[c], method="Radau", jac=jac, rtol=1e-10, atol=1e-22)
        ref = sol.y[:, -1]
        sig = np.abs(ref) > 1e-14
        assert np.abs(out[c] - ref)[sig].max() / np.abs(ref)[sig].max() < 1e-5


def test_f32_literal_switch_changes_results_slightly(oracle):
    var, fix, rc = util.random_cells("gas", 2, 500)
    a = oracle.integrate(0, rc, fix, var, opts=oracle.default_opts(f32_literals=1))[0]
    b = oracle.integrate(0, rc, fix, var, opts=oracle.default_opts(f32_literals=0))[0]
    assert not np.array_equal(a, b)
    assert util.rel_err(a, b, floor=1e-16).max() < 1e-3


def test_threads_do_not_change_results(oracle):
    var, fix, rc = util.random_cells("gas", 40, 600)
    a = oracle.integrate(0, rc, fix, var, nthreads=1)
    b = oracle.integrate(0, rc, fix, var, nthreads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
