"""CPU oracle of SUBROUTINE sedp / sedl / sedc with advsed0 / advsed1 / vterm (oracle/sed_oracle.c restating
str.f90:2257-2864, 5522-5691) - pinned by tests/golden/sed_reference.npz, the result of executing the reference's
own Fortran statements (tests/golden/make_sed_reference.py), and by the properties of the scheme: the advection is
conservative and positive definite, nothing settles through the top, an empty class books the deposition of the
class before it (the reference's x0, str.f90:2352 / 2397)."""
import os

import numpy as np

from mistra_b200 import kon, sed as sm
from oracle import sed_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sed_reference.npz")


def fixture():
    z = np.load(GOLD)
    nf, n, nka, nkt, nkc, j2, j6, j1, nkc_l = [int(x) for x in z["sizes"]]
    g = dict(nka=nka, nkt=nkt, rq=z["rq"], e=z["e"], kw=z["kw"])
    return z, g, nf, nkc_l


def close(a, b, rtol=1e-12):
    """bit-identical unless libm's exp / log entered (vterm of large drops, sedc)"""
    return np.allclose(a, b, rtol=rtol, atol=0.0)


def run_fixture(sedp, sedl, sedc):
    z, g, nf, nkc_l = fixture()
    dt = float(z["dt"])
    ff, dg = sedp(g, dt, nf, z["detw"], z["deta"], z["t"], z["p"], z["vd"], z["ff0"], z["diag0"])
    sl, si = sedl(dt, nf, nkc_l, z["detw"], z["deta"], z["t"], z["p"], z["rc"], z["vt"], z["vdm"], z["sl0"], z["si0"])
    s1 = sedc(dt, z["detw"], z["deta"], z["sedc_vg"], z["sedc_es1"], z["sedc_s0"])
    return z, ff, dg, sl, si, s1


def oracle_sedl(dt, nf, nkc_l, detw, deta, t, p, rc, vt, vdm, sl1=None, sion1=None):
    f = lambda s: None if s is None else so.sedl(dt, nf, nkc_l, detw, deta, t, p, rc, vt, vdm, s)
    return f(sl1), f(sion1)


def test_advection_schemes_and_vterm_against_the_reference_statements():
    z = np.load(GOLD)
    for i in range(len(z["adv_y"])):
        assert np.array_equal(so.advsed(0, z["adv_c"][i], z["adv_y"][i]), z["adv_y0"][i])
        assert np.array_equal(so.advsed(1, z["adv_c"][i], z["adv_y"][i]), z["adv_y1"][i])
    v = so.vterm(z["vterm_a"], z["vterm_t"], z["vterm_p"])
    small = z["vterm_a"] <= 1e-5
    assert small.any() and (~small).any()
    assert np.array_equal(v[small], z["vterm"][small]) and close(v, z["vterm"], 1e-14)


def test_sedp_sedl_sedc_against_the_reference_statements():
    z, ff, dg, sl, si, s1 = run_fixture(so.sedp, oracle_sedl, so.sedc)
    assert close(ff, z["ff1"]) and close(dg, z["diag1"])
    assert close(sl, z["sl1"]) and close(si, z["si1"]) and close(s1, z["sedc_s1"])
    small = z["rq"] <= 10.0                                        # no exp / log below 10 um
    assert np.array_equal(ff[:, :, small], z["ff1"][:, :, small])
    assert (z["ff0"] != z["ff1"]).mean() > 0.3 and (z["sl0"] != z["sl1"]).mean() > 0.3


def test_advsed1_conserves_and_stays_positive():
    r = np.random.default_rng(3)
    for nf in (8, 37, 100):
        y = r.uniform(0, 1, nf) * (r.uniform(size=nf) < 0.7)
        c = -r.uniform(0, 0.9, nf)
        c[0] = c[1]
        y[0] = y[1]
        o = so.advsed(1, c, y)
        assert (o >= 0).all()
        assert abs(o.sum() - y.sum()) <= 1e-13 * y.sum()          # fluxes telescope; y(1) collects what leaves
        assert np.array_equal(so.advsed(1, np.zeros(nf), y), y)    # nothing moves at rest


def test_sedp_properties():
    g = kon.kon_grid()
    d = sm.synthetic_columns(g, 3, seed=4)
    nf, dt = 100, 10.0
    ff, dg = so.sedp(g, dt, nf, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"], d["diag"])
    assert np.array_equal(ff[:, 0], d["ff"][:, 0]) and np.array_equal(ff[:, nf:], d["ff"][:, nf:])    # untouched levels
    assert (ff >= 0).all() and (ff != d["ff"]).any()
    settled = (d["ff"][:, 1:nf] * d["detw"][None, 1:nf, None, None]).sum(axis=1) > 1e-6
    assert settled.any() and not settled.all()
    assert np.array_equal(ff[:, :, ~settled[0]][0], d["ff"][:, :, ~settled[0]][0])                    # skipped classes
    assert np.array_equal(ff[:, nf - 1][settled], ff[:, nf - 2][settled])                             # str.f90:2391
    # column mass of a settled class only leaves through the ground: what is lost is what x0 books
    assert (dg[:, 0] >= 0).all() and (dg[:, 1] >= d["diag"][:, 1]).all()
    # the carried x0: an empty class after a depositing one books that deposition again
    d2 = {k: v[:1].copy() if v.ndim > 1 and k not in ("detw", "deta") else v for k, v in d.items()}
    d2["ff"][:] = 0.0
    q = np.argwhere(g["rq"] > 30.0)[0]
    d2["ff"][0, 1:40, q[0], q[1]] = 5.0
    _, one = so.sedp(g, dt, nf, d2["detw"], d2["deta"], d2["t"], d2["p"], d2["vd"], d2["ff"], np.zeros((1, 4)))
    after = g["nkt"] * g["nka"] - (q[0] * g["nkt"] + q[1])                                            # classes from q on
    x0_q = one[0, 1] / (g["e"][q[1]:].sum() + (g["nka"] - 1 - q[0]) * g["e"].sum()) / d2["detw"][1]
    assert one[0, 1] > 0 and after > 1 and x0_q > 0
    assert np.isclose(one[0, 2] + one[0, 3], one[0, 1], rtol=1e-12) and np.isclose(one[0, 0] * dt, one[0, 1], rtol=1e-12)


def test_sedl_properties():
    g = kon.kon_grid()
    d = sm.synthetic_columns(g, 2, seed=5)
    nf, dt = 100, 10.0
    sl, si = oracle_sedl(dt, nf, 3, d["detw"], d["deta"], d["t"], d["p"], d["rc"], d["vt"], d["vdm"], d["sl1"], d["sion1"])
    for o, a in ((sl, d["sl1"]), (si, d["sion1"])):
        assert np.array_equal(o[:, :, 3], a[:, :, 3]) and np.array_equal(o[:, nf - 1:], a[:, nf - 1:])   # bin 4, upper levels
        assert (o >= 0).all() and (o[:, 1:nf - 1, :3] != a[:, 1:nf - 1, :3]).any()
        # column burden below nf-1 + deposit changes only by what enters from level nf
        gain = o[:, 0, :3] - a[:, 0, :3]
        assert (gain >= 0).all() and (gain > 0).any()
