"""CPU oracle of the condensation step (oracle/kon_oracle.c restating subkon + advec,
str.f90:4987-5204, 5321-5516) - pinned by the scheme's own guarantees and the committed
golden layers (the reference has no fixtures: "parity unpinned")."""
import os

import numpy as np
import pytest

from mistra_b200 import kon
from oracle import kon_oracle as ko

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kon_layers.npz")
R0 = 8.3144743 / 28.96546e-3


@pytest.fixture(scope="module")
def grid():
    return kon.kon_grid()


def upstream_u(c):
    """u(jt) of str.f90:5170-5175 from bin-interface growth rates c."""
    n = len(c)
    u = np.empty(n)
    u[0] = max(0.0, c[0])
    u[1:n - 1] = 0.5 * (c[1:n - 1] + np.abs(c[1:n - 1]) + c[0:n - 2] - np.abs(c[0:n - 2]))
    u[n - 1] = min(0.0, c[n - 2])
    return u


def test_advec_identity_shift_and_conservation():
    r = np.random.default_rng(0)
    n = 70
    y = np.where(r.uniform(size=n) < 0.6, 10.0 ** r.uniform(-6, 3, n), 0.0)
    out, err = ko.advec(10.0, np.zeros(n), y)
    assert err == 0 and np.array_equal(out, y)                           # u = 0: nothing moves
    # Courant number exactly +1 / -2: pure shifts by whole bins (c0 = 0 branch)
    y2 = y.copy(); y2[-3:] = 0.0
    out, err = ko.advec(10.0, np.full(n, 0.1), y2)
    assert err == 0 and np.array_equal(out[1:], y2[:-1]) and out[0] == 0.0
    y3 = y.copy(); y3[:3] = 0.0
    out, err = ko.advec(10.0, np.full(n, -0.2), y3)
    assert err == 0 and np.array_equal(out[:-2], y3[2:])
    # upstream velocities as subkon builds them: positive definite and conservative
    for s in range(20):
        c = r.uniform(-0.25, 0.25, n) * r.choice([0.05, 1.0, 3.0])
        u = upstream_u(c)
        out, err = ko.advec(10.0, u, y)
        assert err == 0 and (out >= 0).all()
        assert np.isclose(out.sum(), y.sum(), rtol=1e-13)
    # the reference aborts when a particle leaves the grid: reported, not fatal
    u = np.full(n, 0.1); yb = np.zeros(n); yb[-1] = 1.0
    out, err = ko.advec(10.0, u * 0.5, yb)
    assert err == 1


def test_advec_fractional_courant_moves_the_centre_of_mass():
    n = 70
    x = np.arange(n)
    y = np.exp(-0.5 * ((x - 30.0) / 4.0) ** 2)
    y[y < 1e-12] = 0.0                                                    # nothing near the grid ends
    for cfl in (0.3, 0.5, 1.7):
        out, err = ko.advec(10.0, np.full(n, cfl / 10.0), y)
        assert err == 0
        shift = (out * x).sum() / out.sum() - (y * x).sum() / y.sum()
        assert abs(shift - cfl) < 2e-3                                    # 4th-order area-preserving fluxes


def test_subkon_water_and_heat_budget(grid):
    d = kon.synthetic_layers(grid, 96, seed=3)
    ffk, to, xm1o, st = ko.subkon(grid, 10.0, d["ffk"], d["totr"], d["dfdt"], d["feualt"], d["pp"], d["to"], d["tn"],
                                  d["xm1o"], d["xm1n"], d["kr"])
    assert (st >= 1).all() and (st <= 10).all()                          # every layer converged
    assert (ffk >= 0).all()
    assert np.allclose(ffk.sum(axis=2), d["ffk"].sum(axis=2), rtol=1e-12, atol=1e-300)   # particles per dry class
    dw = ((ffk - d["ffk"]) * grid["e"][None, None, :]).sum(axis=(1, 2))
    rho = d["pp"] / (R0 * d["to"] * (1.0 + 0.61 * d["xm1o"]))
    assert np.allclose(d["xm1n"] - xm1o, dw / rho, rtol=1e-9, atol=1e-18)         # vapour lost = liquid gained
    xldcp = (3138708.0 - 2339.4 * d["to"]) / 1005.0
    assert np.allclose(to - d["tn"], xldcp * dw / rho, rtol=1e-9, atol=1e-12)     # latent heating (to ~ 280 K: ulp 6e-14)
    # supersaturated layers condense, clearly sub-saturated ones evaporate
    assert (dw[d["feualt"] > 1.002] > 0).all()


def test_subkon_empty_layer_and_golden(grid):
    d = kon.synthetic_layers(grid, 2, seed=4)
    d["ffk"][:] = 0.0
    ffk, to, xm1o, st = ko.subkon(grid, 10.0, d["ffk"], d["totr"], d["dfdt"], d["feualt"], d["pp"], d["to"], d["tn"],
                                  d["xm1o"], d["xm1n"], d["kr"])
    assert not ffk.any() and np.array_equal(to, d["tn"]) and np.array_equal(xm1o, d["xm1n"])
    gd = np.load(GOLD)
    ffk, to, xm1o, st = ko.subkon(grid, float(gd["dt"]), gd["ffk"], gd["totr"], gd["dfdt"], gd["feualt"], gd["pp"],
                                  gd["to"], gd["tn"], gd["xm1o"], gd["xm1n"], gd["kr"])
    assert np.array_equal(st, gd["status"])
    assert np.allclose(ffk, gd["ffk_out"], rtol=1e-12, atol=1e-14 * gd["ffk_out"].max())
    assert np.allclose(to, gd["to_out"], rtol=1e-14) and np.allclose(xm1o, gd["xm1o_out"], rtol=1e-13)


def test_layers_dry_branch_is_koehler_equilibrium(grid):
    """kon's layer loop (str.f90:4615-4772): dry layers (feu < 0.7) go through equil."""
    st = kon.synthetic_columns(grid, 64, seed=8, dry_fraction=0.5)
    o = ko.layers(grid, 10.0, True, st)
    dry = st["feu"] < 0.7
    assert (o["status"][dry] == 0).all() and (o["status"][~dry] >= 1).all()
    ff = o["ff"]
    # every dry class sits in exactly one water bin and keeps its particles
    assert np.allclose(ff.sum(axis=2), st["ff"].sum(axis=2), rtol=1e-13, atol=1e-300)
    occupied = (ff[dry] > 0).sum(axis=2)
    assert (occupied <= 1).all()
    # that bin brackets the equilibrium water mass: ew(jt-1) < eg <= ew(jt), with rg from the
    # Koehler equation  ln(feu) = a0/rg - b0*rn^3/(rg^3 - rn^3)
    k = np.nonzero(dry)[0][0]
    feun = st["xm1"][k] * st["p"][k] / ((0.62198 + 0.37802 * st["xm1"][k]) * kon.p21(st["t"][k]))
    assert np.isclose(o["feu"][k], feun, rtol=1e-14)
    a0 = grid["a0m"] / st["t"][k]
    for ia in (5, 30, 60):
        jt = int(np.argmax(ff[k, ia] > 0))
        lo = grid["ew"][jt - 1] if jt > 0 else 0.0
        # solve the Koehler equation independently by bisection on the water mass
        rn = grid["rn"][ia]
        f = lambda rg: np.log(feun) - (a0 / rg - 2.0 * grid["b0m"][ia] * rn ** 3 / (rg ** 3 - rn ** 3))
        a, b = rn * (1 + 1e-9), rn * 50
        for _ in range(200):
            m = 0.5 * (a + b)
            a, b = (m, b) if f(m) > 0 else (a, m)          # f falls from +inf (rg -> rn) through 0
        eg = 4.0e-9 * np.pi / 3.0 * (a ** 3 - rn ** 3)
        assert lo * (1 - 1e-6) <= eg <= grid["ew"][jt] * (1 + 1e-6)
    # untouched in the dry branch: temperature and vapour; dtcon = 0
    assert np.array_equal(o["t"][dry], st["t"][dry]) and np.array_equal(o["xm1"][dry], st["xm1"][dry])
    assert not o["dtcon"][dry].any()
    # humid layers are exactly subkon + the write-back of str.f90:4708-4721
    i = np.nonzero(~dry)[0]
    f2, to, xm1o, s2 = ko.subkon(grid, 10.0, st["ff"][i], st["totrad"][i], st["dfddt"][i], st["feu"][i], st["p"][i],
                                 st["talt"][i], st["t"][i], st["xm1a"][i], st["xm1"][i], st["nar"][i])
    assert np.array_equal(f2, ff[i]) and np.array_equal(to, o["t"][i]) and np.array_equal(to, o["talt"][i])
    assert np.array_equal(xm1o, o["xm1"][i]) and np.array_equal(s2, o["status"][i])
    assert np.allclose(o["dtcon"][i], (to - st["t"][i]) / 10.0, rtol=0, atol=0)
    assert np.allclose(o["xm2"], (ff * grid["e"][None, None, :]).sum(axis=(1, 2)), rtol=1e-12)
    # bin sums for konc
    kw, ka = grid["kw"], grid["ka"]
    aer = np.arange(grid["nkt"])[None, :] < kw[:, None]
    assert np.allclose(o["part_o_a"], (st["ff"] * aer[None]).sum(axis=2), rtol=1e-13, atol=1e-300)
    assert np.allclose(o["part_n_d"], (ff * ~aer[None]).sum(axis=2), rtol=1e-13, atol=1e-300)
    assert np.allclose(o["pntot"][:, 0], o["part_n_a"][:, :ka].sum(axis=1), rtol=1e-13)
    assert np.allclose(o["vol2"][:, 3], o["vol1_d"][:, ka:].sum(axis=1), rtol=1e-13)
