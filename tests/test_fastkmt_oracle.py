"""CPU oracle of SUBROUTINE fast_k_mt_a / fast_k_mt_t and FUNCTION vterm (oracle/fastkmt_oracle.c
restating kpp.f90:2683-2947 and str.f90:2793-2864) - pinned by a numpy evaluation of the same
integrals, the write rules (cm switch, cw > 0, untouched entries) and golden layers (the reference has
no fixtures: "parity unpinned")."""
import os

import numpy as np

from mistra_b200 import fastkmt, kon
from oracle import cwrc_oracle as cwo
from oracle import fastkmt_oracle as fko

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastkmt_layers.npz")
NSPEC = {"aer": 262, "tot": 424}


def inputs(n, seed, mech="aer", dense=0.5, g=None):
    """Synthetic layers: spectra from the condensation generator with a random share of the other grid
    points populated, cw / cm from the cw_rc oracle (so some bins are without chemistry), accommodation
    coefficients with zeros, mean molecular speeds of 100-700 m/s."""
    g = g or kon.kon_grid()
    d = kon.synthetic_columns(g, n, seed=seed, dry_fraction=0.4)
    r = np.random.default_rng(seed)
    ff = d["ff"] * 10.0 ** r.uniform(-1, 3, (n, 1, 1))
    extra = (r.uniform(size=ff.shape) < dense * r.uniform(size=(n, 1, 1))) * 10.0 ** r.uniform(-6, 0, ff.shape)
    ff = ff + extra
    feu = np.where(r.uniform(size=n) < 0.15, r.uniform(0.2, 0.45, n), d["feu"])
    cloud = r.uniform(size=(n, 4)) < 0.5
    _, cw, cm, _ = cwo.cw_rc(g, ff, feu, cloud)
    t, p = d["t"], d["p"]
    freep = 2.28e-5 * t / p                                         # liq_parm, kpp.f90:585
    ns = NSPEC[mech]
    alpha = np.where(r.uniform(size=(n, ns)) < 0.2, 0.0, 10.0 ** r.uniform(-4, 0, (n, ns)))
    vmean = r.uniform(100.0, 700.0, (n, ns))
    xkmt0 = r.uniform(1.0, 2.0, (n, 4, ns))                          # previous values: must survive where not assigned
    vt0 = r.uniform(1.0, 2.0, (n, 4))
    return dict(g=g, lex=fastkmt.lex(mech), ff=ff, freep=freep, t=t, p=p, cw=cw, cm=cm, alpha=alpha, vmean=vmean,
                xkmt=xkmt0, vt=vt0)


def masks(g, nkc_l=4, ial=1):
    nka, nkt, ka = g["nka"], g["nkt"], g["ka"]
    aer = np.arange(nkt)[None, :] < np.asarray(g["kw"])[:, None]
    small = (np.arange(nka) < ka)[:, None]
    first = (np.arange(nka) >= ial - 1)[:, None]
    m = [aer & small & first, aer & ~small, ~aer & small & first, ~aer & ~small]
    return m[:nkc_l]


def vterm_numpy(a, t, p):
    gg, r0 = 9.80665, 8.3144743 / 28.96546e-3
    rho_a = p / (r0 * t)
    eta = 3.7957e-06 + 4.9e-08 * t
    stokes = 2 * gg / 9 * a * a * (1000.0 - rho_a) / eta * (1 + 1.26 * 6.6e-8 * 101325 / 293.15 * t / (a * p))
    x = np.log(32 * gg / 3 * a ** 3 * (1000.0 - rho_a) * rho_a / eta ** 2)
    y = np.polyval([-.327815e-5, .855176e-4, -.578878e-3, -.987059e-3, -.153193e-2, .992696, -.318657e1], x)
    return np.where(a <= 1e-5, stokes, eta * np.exp(y) / (2 * rho_a * a))


def test_vterm():
    a = 10.0 ** np.linspace(-8, np.log10(5e-4), 400)      # the polynomial regime ends at 535 um (no third regime)
    v = fko.vterm(a, 285.0, 9.5e4)
    assert np.allclose(v, vterm_numpy(a, 285.0, 9.5e4), rtol=1e-12)
    assert np.all(np.diff(v) > 0) and 1e-8 < v[0] < 1e-6 and 3 < v[-1] < 5       # 10 nm ... 0.5 mm drops
    i = np.searchsorted(a, 1e-5)
    assert abs(v[i] / v[i - 1] - 1) < 0.1                                         # regimes join


def check_against_numpy(x, nkc_l, ial, xk, vt):
    g, lex = x["g"], x["lex"]
    rqm = np.asarray(g["rq"]) * 1e-6
    z4pi3 = 4.0 * 3.1415926535897932 / 3.0
    n = x["ff"].shape[0]
    q = rqm[None] / x["freep"][:, None, None]
    vts = fko.vterm(rqm[None], x["t"][:, None, None], x["p"][:, None, None])
    touched = np.zeros(xk.shape, dtype=bool)
    for kc, m in enumerate(masks(g, nkc_l, ial)):
        cw, on = x["cw"][:, kc], x["cm"][:, kc] > 0
        pos = cw > 0
        s = (rqm[None] ** 3 * vts * x["ff"] * 1e6 * m).sum(axis=(1, 2))
        assert np.allclose(vt[pos, kc], z4pi3 / cw[pos] * s[pos], rtol=1e-11)
        assert np.array_equal(vt[~pos, kc], x["vt"][~pos, kc])
        for l, sp in enumerate(lex - 1):
            al = x["alpha"][:, sp]
            x1 = np.where(al > 0, 4.0 / (3.0 * np.where(al > 0, al, 1.0)), 0.0)
            s = (x["vmean"][:, sp, None, None] / (q + x1[:, None, None]) * rqm[None] ** 2 * x["ff"] * 1e6 * m).sum(axis=(1, 2))
            w = on & pos
            assert np.allclose(xk[w, kc, sp], z4pi3 / cw[w] * s[w], rtol=1e-11)
            touched[w, kc, sp] = True
    assert np.array_equal(xk[~touched], x["xkmt"][~touched])                      # everything else keeps its value
    for kc in range(nkc_l, 4):
        assert np.array_equal(vt[:, kc], x["vt"][:, kc])
    return touched


def call(fn, x, **kw):
    return fn(x["g"], x["lex"], x["ff"], x["freep"], x["t"], x["p"], x["cw"], x["cm"], x["alpha"], x["vmean"],
              x["xkmt"], x["vt"], **kw)


def test_against_numpy_aer_and_tot():
    x = inputs(120, 3, "aer")
    xk, vt = call(fko.fast_k_mt, x, nkc_l=4)
    touched = check_against_numpy(x, 4, 1, xk, vt)
    on = x["cm"] > 0
    assert 0.05 < on[:, :2].mean() < 0.98 and on[:, 2:].any() and touched.any()
    xk2, vt2 = call(fko.fast_k_mt, x, nkc_l=2, ial=2)                             # nkc_l = 2, ifeed = 2
    check_against_numpy(x, 2, 2, xk2, vt2)
    y = inputs(60, 4, "tot")
    xk, vt = call(fko.fast_k_mt, y)
    check_against_numpy(y, 4, 1, xk, vt)


def test_golden_layers():
    z = np.load(GOLD)
    x = {k: z[k] for k in ("lex", "ff", "freep", "t", "p", "cw", "cm", "alpha", "vmean", "xkmt", "vt")}
    x["g"] = kon.kon_grid()
    xk, vt = call(fko.fast_k_mt, x)
    assert np.allclose(xk, z["xkmt_out"], rtol=1e-15, atol=0) and np.allclose(vt, z["vt_out"], rtol=2e-15, atol=0)
