"""SUBROUTINE sedp / sedl / sedc on the device (row N4, the rest; include/mistra_sed.h) vs the fixtures made from the
reference's own statements and vs the CPU oracle (-m gpu).  Every profile is one sequential recurrence in the
reference's operation order: bit for bit wherever no exp / log enters (advsed0 / advsed1, vterm below 10 um), 1e-12
where CUDA's exp / log do (vterm of large drops, sedc)."""
import numpy as np
import pytest

from mistra_b200 import kon, sed as sm
from oracle import sed_oracle as so
from tests.test_sed_oracle import close, oracle_sedl, run_fixture

pytestmark = pytest.mark.gpu


def test_reference_statement_fixtures(cuda_device, kpp):
    n0 = sm.launch_count()
    z, ff, dg, sl, si, s1 = run_fixture(sm.sedp, sm.sedl, sm.sedc)
    assert sm.launch_count() == n0 + 5 + 2 + 1                       # sedp (scan, list, units, work, diagnostics), sl1 + sion1, sedc
    assert close(ff, z["ff1"]) and close(dg, z["diag1"])
    small = z["rq"] <= 10.0
    assert np.array_equal(ff[:, :, small], z["ff1"][:, :, small])
    assert close(sl, z["sl1"]) and close(si, z["si1"]) and close(s1, z["sedc_s1"])


def test_divisions_by_constants_equal_the_ieee_division(cuda_device, kpp):
    """advsed1's /24, /48, /1920, /384, /768, /3840 as multiply + two FMAs: 6 x 2^28 comparisons with the division on the
    device, incl. denormals, Inf / NaN, exactly divisible and all-ones significands."""
    for seed in (1, 2):
        assert sm.divc_selftest(1 << 28, seed) == 0
    assert sm.divc_selftest(0) == 0


@pytest.mark.parametrize("ncol,seed,n,nf,dt", [(6, 11, 150, 100, 10.0), (3, 12, 150, 100, 60.0), (2, 13, 40, 33, 10.0)])
def test_sedp_vs_oracle(cuda_device, kpp, ncol, seed, n, nf, dt):
    g = kon.kon_grid()
    d = sm.synthetic_columns(g, ncol, n=n, nf=nf, seed=seed)
    a = (g, dt, nf, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"], d["diag"])
    ff_r, dg_r = so.sedp(*a)
    ff, dg = sm.sedp(*a)
    assert close(ff, ff_r) and close(dg, dg_r)
    small = g["rq"] <= 10.0
    assert np.array_equal(ff[:, :, small], ff_r[:, :, small]) and (ff != d["ff"]).any()
    assert np.array_equal(ff == d["ff"], ff_r == d["ff"])            # the same classes settle


@pytest.mark.parametrize("ncol,seed,nkc_l,dt", [(5, 21, 4, 10.0), (3, 22, 2, 60.0)])
def test_sedl_vs_oracle(cuda_device, kpp, ncol, seed, nkc_l, dt):
    g = kon.kon_grid()
    d = sm.synthetic_columns(g, ncol, seed=seed)
    a = (dt, 100, nkc_l, d["detw"], d["deta"], d["t"], d["p"], d["rc"], d["vt"], d["vdm"])
    ref = oracle_sedl(*a, d["sl1"], d["sion1"])
    out = sm.sedl(*a, d["sl1"], d["sion1"])
    for o, r in zip(out, ref):
        assert close(o, r)
    fine = d["rc"][..., :nkc_l].max() <= 1e-5                        # every bin below 10 um: no exp / log anywhere
    if fine:
        assert all(np.array_equal(o, r) for o, r in zip(out, ref))
    only = sm.sedl(*a, None, d["sion1"])
    assert only[0] is None and np.array_equal(only[1], out[1])


def test_device_entries_and_edges(cuda_device, kpp):
    import torch
    g = kon.kon_grid()
    d = sm.synthetic_columns(g, 4, seed=31)
    t = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)
    gd = dict(nka=g["nka"], nkt=g["nkt"], rq=t(g["rq"]), e=t(g["e"]), kw=t(g["kw"], np.int32))
    dv = {k: t(v) for k, v in d.items()}
    sm.sedp_device(gd, 10.0, 100, dv["detw"], dv["deta"], dv["t"], dv["p"], dv["vd"], dv["ff"], dv["diag"])
    sm.sedl_device(10.0, 100, 4, dv["detw"], dv["deta"], dv["t"], dv["p"], dv["rc"], dv["vt"], dv["vdm"], dv["sl1"], dv["sion1"])
    torch.cuda.synchronize()
    ff, dg = sm.sedp(g, 10.0, 100, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"], d["diag"])
    sl, si = sm.sedl(10.0, 100, 4, d["detw"], d["deta"], d["t"], d["p"], d["rc"], d["vt"], d["vdm"], d["sl1"], d["sion1"])
    assert np.array_equal(dv["ff"].cpu().numpy(), ff) and np.array_equal(dv["diag"].cpu().numpy(), dg)
    assert np.array_equal(dv["sl1"].cpu().numpy(), sl) and np.array_equal(dv["sion1"].cpu().numpy(), si)
    r = np.random.default_rng(1)
    vg, es1, s1 = np.where(r.uniform(size=93) < 0.5, 0.0, 10.0 ** r.uniform(-6, -2, 93)), 10.0 ** r.uniform(8, 12, 93), \
        10.0 ** r.uniform(-12, -7, (4, 150, 93))
    s1d = t(s1)
    sm.sedc_device(10.0, dv["detw"], dv["deta"], t(vg), t(es1), s1d)
    torch.cuda.synchronize()
    ref = so.sedc(10.0, d["detw"], d["deta"], vg, es1, s1)
    assert close(s1d.cpu().numpy(), ref) and np.array_equal(s1d.cpu().numpy()[:, 2:], s1[:, 2:])
    # empty ensemble; bad sizes are refused
    e = {k: (v[:0] if k not in ("detw", "deta") else v) for k, v in d.items()}
    assert sm.sedp(g, 10.0, 100, e["detw"], e["deta"], e["t"], e["p"], e["vd"], e["ff"], e["diag"])[0].shape == (0, 150, 70, 70)
    with pytest.raises(ValueError):
        sm.sedp(g, 10.0, 100, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"][:, :, :5], d["diag"])
    with pytest.raises(Exception):
        sm.sedp(g, 10.0, 200, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"], d["diag"])      # nf > n


def test_full_size_properties(cuda_device, kpp):
    """48 columns x 150 levels x 70 x 70 classes on the device: positivity, untouched levels and classes, and the
    budget of every settled class - what left the column is what the diagnostics book (sum over classes of
    x0 * e * detw(2) with x0 the flux through the ground)."""
    g = kon.kon_grid()
    d = sm.synthetic_columns(g, 48, seed=41)
    nf, dt = 100, 10.0
    ff, dg = sm.sedp(g, dt, nf, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"], np.zeros_like(d["diag"]))
    assert (ff >= 0).all() and np.array_equal(ff[:, nf:], d["ff"][:, nf:]) and np.array_equal(ff[:, 0], d["ff"][:, 0])
    w = d["detw"][None, 1:nf - 1, None, None]
    settled = (d["ff"][:, 1:nf] * d["detw"][None, 1:nf, None, None]).sum(axis=1) > 1e-6
    lost = ((d["ff"][:, 1:nf - 1] - ff[:, 1:nf - 1]) * w).sum(axis=1)                  # per class, levels 2..nf-1
    inflow_free = d["ff"][:, nf - 2] == 0.0                                            # nothing enters from level nf
    ok = settled & inflow_free[:, :, :] & (d["ff"][:, nf - 1] == 0.0)
    assert ok.any()
    assert (lost[ok] >= -1e-9 * np.abs(d["ff"]).max()).all()
    assert np.allclose(dg[:, 2] + dg[:, 3], dg[:, 1], rtol=1e-12) and np.allclose(dg[:, 0] * dt, dg[:, 1], rtol=1e-12)
