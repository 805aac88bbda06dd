"""CPU oracle of the 2-D bin redistribution (oracle/bins_oracle.c, restating
str.f90:5916-6134) - pinned by properties the algorithm guarantees and by the committed
golden layers (the reference has no fixtures of its own: "parity unpinned")."""
import os

import numpy as np
import pytest

from mistra_b200 import bins
from oracle import bins_oracle as bo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bins_layers.npz")


@pytest.fixture(scope="module")
def grid():
    return bins.particle_grid()


@pytest.fixture(scope="module")
def case(grid):
    d = bins.synthetic_layers(grid, 48, seed=5)
    sap, smp, so = bo.snapshot(grid, d["ff"], d["cm"], d["sion1"])
    out = bo.redistribute(grid, d["ff"], d["cm"], d["cw"], sap, smp, so, d["sion1_new"], d["sl1"])
    return d, sap, smp, so, out


def test_particle_grid_matches_reference_construction(grid):
    # str.f90:1653-1705: geometric mass grids; class radius of the mean mass; ka, kw limits
    en, rn = grid["en"], grid["rn"]
    assert np.allclose(en[1:] / en[:-1], en[1] / en[0], rtol=1e-12)
    assert np.allclose(rn, (en / (4.0 / 3.0 * np.pi * 2000.0)) ** (1.0 / 3.0) * 1.0e4, rtol=1e-12)
    assert rn[grid["ka"] - 1] <= 0.5 < rn[grid["ka"]]                   # ka = last class with rn <= 0.5 um
    assert (np.diff(grid["kw"]) >= 0).all() and grid["kw"][-1] == grid["nkt"]
    assert grid["rq"].shape == (70, 70) and (grid["rq"] > rn[:, None]).all()


def test_snapshot_sums(grid, case):
    d, sap, smp, so, _ = case
    ff, kw, ka, en = d["ff"], grid["kw"], grid["ka"], grid["en"]
    jt = np.arange(grid["nkt"])[None, :]
    aer = jt < kw[:, None]
    small = (np.arange(grid["nka"]) < ka)[:, None]
    last = (np.arange(grid["nka"]) == grid["nka"] - 1)[:, None]        # ia = nka is never in bins 2/4
    for kc, mk in enumerate([small & aer, ~small & aer & ~last, small & ~aer, ~small & ~aer & ~last]):
        on = d["cm"][:, kc] != 0
        assert np.allclose(sap[on, kc], (ff * mk[None]).sum(axis=(1, 2))[on], rtol=1e-12)
        assert np.allclose(smp[on, kc], (ff * (en[:, None] * mk)[None]).sum(axis=(1, 2))[on], rtol=1e-12)
        assert (sap[~on, kc] == 0).all() and (smp[~on, kc] == 0).all()
    for i, l in enumerate(bins.LJ2):
        on = d["cm"] != 0
        assert np.array_equal(so[:, :, i][on], d["sion1"][:, :, l - 1][on])


def test_number_per_water_bin_is_conserved(case):
    d, *_, out = case
    ff2 = out[0]
    n0, n1 = d["ff"].sum(axis=1), ff2.sum(axis=1)                        # sum over dry classes
    assert np.allclose(n1, n0, rtol=1e-13, atol=1e-300) and (ff2 >= 0).all()


def test_dry_mass_changes_by_the_ion_mass_change(grid, case):
    # the linear split over (ix, ix+1) conserves x0 = en*(1 + den*sap/smp) per particle, so the
    # bin's dry mass moves by exactly den*sap (str.f90:6023-6060)
    d, sap, smp, so, out = case
    ff2 = out[0]
    en = grid["en"]
    dm = ((ff2 - d["ff"]) * en[None, :, None]).sum(axis=(1, 2))
    ds = np.stack([(d["sion1_new"][:, :, l - 1] - so[:, :, i]) for i, l in enumerate(bins.LJ2)], axis=-1)
    den_sap = (ds * np.array(bins.ION_MASS)).sum(axis=-1) * 1.0e-6 * 1000.0   # = den * sap  [mg cm^-3]
    on = (d["cm"] != 0) & (sap > 1e-6)
    expect = (den_sap * on).sum(axis=1)
    # (up to the clipping of x0 at the ends of the dry-mass grid, str.f90:6037-6043)
    assert (np.abs(dm - expect) <= 5e-3 * np.abs(den_sap * on).sum(axis=1)).all()


def test_dissolved_species_are_conserved_across_bins(case):
    d, *_, out = case
    _, si2, sl2, nwarn = out
    assert np.allclose(sl2.sum(axis=1), d["sl1"].sum(axis=1), rtol=1e-13)
    assert np.allclose(si2.sum(axis=1), d["sion1_new"].sum(axis=1), rtol=1e-13)
    assert (nwarn == 0).all()
    assert (np.abs(sl2 - d["sl1"]).max(axis=(1, 2)) > 0).mean() > 0.5   # volume did cross bin limits


def test_identity_without_mass_change_and_for_dry_layers(grid):
    d = bins.synthetic_layers(grid, 6, seed=9)
    sap, smp, so = bo.snapshot(grid, d["ff"], d["cm"], d["sion1"])
    ff2, si2, sl2, _ = bo.redistribute(grid, d["ff"], d["cm"], d["cw"], sap, smp, so, d["sion1"], d["sl1"])
    assert np.array_equal(ff2, d["ff"]) and np.array_equal(si2, d["sion1"]) and np.array_equal(sl2, d["sl1"])
    cm0 = np.zeros_like(d["cm"])
    sap0, smp0, _ = bo.snapshot(grid, d["ff"], cm0, d["sion1"])
    assert not sap0.any() and not smp0.any()
    ff3, si3, sl3, _ = bo.redistribute(grid, d["ff"], cm0, d["cw"], sap0, smp0, so, d["sion1_new"], d["sl1"])
    assert np.array_equal(ff3, d["ff"]) and np.array_equal(si3, d["sion1_new"]) and np.array_equal(sl3, d["sl1"])


def test_growth_moves_up_and_loss_moves_down(grid):
    d = bins.synthetic_layers(grid, 8, seed=11)
    sap, smp, so = bo.snapshot(grid, d["ff"], d["cm"], d["sion1"])
    en = grid["en"]
    for sign in (+1.0, -1.0):
        new = d["sion1"] * (1.0 + sign * 0.3)
        ff2, *_ = bo.redistribute(grid, d["ff"], d["cm"], d["cw"], sap, smp, so, new, d["sl1"])
        mean0 = (d["ff"] * en[None, :, None]).sum(axis=(1, 2)) / d["ff"].sum(axis=(1, 2))
        mean1 = (ff2 * en[None, :, None]).sum(axis=(1, 2)) / ff2.sum(axis=(1, 2))
        assert (sign * (mean1 - mean0) > 0).all()


def test_golden_layers():
    g = np.load(GOLD)
    grid = bins.particle_grid(*g["grid_args"])
    sap, smp, so = bo.snapshot(grid, g["ff"], g["cm"], g["sion1"])
    assert np.array_equal(sap, g["sap"]) and np.array_equal(smp, g["smp"]) and np.array_equal(so, g["sion1o"])
    ff2, si2, sl2, nw = bo.redistribute(grid, g["ff"], g["cm"], g["cw"], sap, smp, so, g["sion1_new"], g["sl1"])
    assert np.array_equal(ff2, g["ff_out"]) and np.array_equal(si2, g["sion1_out"])
    assert np.array_equal(sl2, g["sl1_out"]) and np.array_equal(nw, g["nwarn"])
    assert np.array_equal(ff2[3], g["ff"][3]) and np.array_equal(ff2[5], g["ff"][5])
