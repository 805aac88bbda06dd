"""CPU oracle of SUBROUTINE difc (oracle/difc_oracle.c restating str.f90:3271-3445) - pinned by an
independent banded solve of the same tridiagonal systems (scipy), the invariance of a well-mixed
profile, the untouched levels / bins, and golden columns (the reference has no fixtures: "parity
unpinned")."""
import os

import numpy as np
from scipy.linalg import solve_banded

from mistra_b200 import difc as dm
from oracle import difc_oracle as dfo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "difc_columns.npz")
J1, J5, J2, J6, NKC = 93, 24, 121, 55, 4          # gas_common / global_params sizes of the reference


def inputs(ncol, seed, n=150, nkc_l=4, sizes=(J1, J5, J2 * NKC, J6 * NKC)):
    c = dm.synthetic_columns(ncol, n, seed)
    r = np.random.default_rng(seed + 1)
    fields = []
    for i, row in enumerate(sizes):
        a = 10.0 ** r.uniform(-3, 3, (ncol, 1, row)) * c["am3"][:, :, None] * r.uniform(0.2, 1.8, (ncol, n, row))
        a[r.uniform(size=a.shape) < 0.1] = 0.0
        nproc = row if i < 2 else row // NKC * nkc_l
        fields.append((a, nproc))
    return c, fields


def run(fn, dt, c, fields):
    return fn(dt, c["atkh"], c["w"], c["am3"], c["detw"], c["deta"], fields)


def banded_reference(dt, c, s, col):
    """The system the recurrences of str.f90:3334-3353 solve, for one column and all species at once:
    -xc(k) x(k-1) + (1 + xa(k) + xc(k)) x(k) - xa(k) x(k+1) = x_old(k), k = 2..nm, with x = s / am3,
    x(1) := x_old(2) and x(n) fixed; then the subsidence step."""
    n = c["atkh"].shape[1]
    nm = n - 1
    detw, deta, atkh, am3, w = c["detw"], c["deta"], c["atkh"][col], c["am3"][col], c["w"][col]
    xa = atkh * dt / (detw * deta)
    xc = np.zeros(n); xc[1:] = xa[:-1] * detw[:-1] / detw[1:]
    x = s / am3[:, None]
    m = nm - 1                                              # unknowns: levels index 1 .. nm-1
    ab = np.zeros((3, m))
    ab[1] = 1.0 + xa[1:nm] + xc[1:nm]
    ab[0, 1:] = -xa[1:nm - 1]
    ab[2, :-1] = -xc[2:nm]
    rhs = x[1:nm].copy()
    rhs[0] += xc[1] * x[1]                                  # x(1) := x_old(2)
    rhs[-1] += xa[nm - 1] * x[nm]
    xn = x.copy()
    xn[1:nm] = solve_banded((1, 1), ab, rhs)
    sd = xn * am3[:, None]
    out = sd.copy()
    cc = w * dt / deta
    out[1:nm] = sd[1:nm] - cc[1:nm, None] * (sd[2:nm + 1] - sd[1:nm])
    return out


def test_against_banded_solve_and_untouched_parts():
    c, fields = inputs(6, 3, nkc_l=2)
    outs = run(dfo.difc, 60.0, c, fields)
    for (a, nproc), o in zip(fields, outs):
        assert np.array_equal(o[:, 0], a[:, 0]) and np.array_equal(o[:, -1], a[:, -1])     # levels 1 and n
        assert np.array_equal(o[:, :, nproc:], a[:, :, nproc:])                            # bins kc > nkc_l
        for col in range(a.shape[0]):
            ref = banded_reference(60.0, c, a[col, :, :nproc], col)
            assert np.allclose(o[col, :, :nproc], ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())
    assert any((o != a).any() for (a, _), o in zip(fields, outs))


def test_well_mixed_profile_is_invariant_without_subsidence():
    c, _ = inputs(4, 5)
    c["w"][:] = 0.0
    s = np.repeat(c["am3"][:, :, None], 7, axis=2) * np.arange(1.0, 8.0)
    out = run(dfo.difc, 10.0, c, [(s, 7)])[0]
    assert np.allclose(out, s, rtol=1e-13)


def test_golden_columns():
    z = np.load(GOLD)
    c = {k: z[k] for k in ("atkh", "w", "am3", "detw", "deta")}
    outs = run(dfo.difc, float(z["dt"]), c, [(z["f%d" % i], int(z["nproc"][i])) for i in range(4)])
    for i, o in enumerate(outs):
        assert np.array_equal(o, z["o%d" % i])


# ---- difp (str.f90:3137-3265): the same exchange on ff / rho ----
def difp_inputs(ncol, seed, n=150, row=70 * 70):
    c = dm.synthetic_columns(ncol, n, seed)
    r = np.random.default_rng(seed + 2)
    rho = 1.2 * np.exp(-np.cumsum(c["detw"])[None] / 8000.0) * r.uniform(0.97, 1.03, (ncol, n))
    ff = 10.0 ** r.uniform(-3, 3, (ncol, n, row)) * (r.uniform(size=(ncol, n, row)) < 0.3)
    fsum = r.uniform(1.0, 2.0, (ncol, n))
    return c, rho, ff, fsum


def test_difp_against_difc_recurrence_and_fsum():
    c, rho, ff, fsum = difp_inputs(3, 4, row=60)
    out, fs = dfo.difp(60.0, c["atkh"], c["w"], rho, c["detw"], c["deta"], ff, fsum)
    # the same system as difc with am3 := rho (different rounding order only)
    ref = dfo.difc(60.0, c["atkh"], c["w"], rho, c["detw"], c["deta"], [(ff, 60)])[0]
    assert np.allclose(out[:, 1:-1], ref[:, 1:-1], rtol=1e-12, atol=1e-300)
    assert np.array_equal(out[:, 0], ff[:, 0])                                   # level 1 untouched
    assert np.array_equal(out[:, -1], ff[:, -1] / rho[:, -1, None] * rho[:, -1, None])   # level n: / rho * rho
    assert np.allclose(fs[:, 1:], out[:, 1:].sum(axis=2), rtol=1e-13) and np.array_equal(fs[:, 0], fsum[:, 0])
    assert (out[:, 1:-1] != ff[:, 1:-1]).any()
