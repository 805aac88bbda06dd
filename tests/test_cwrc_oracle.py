"""CPU oracle of SUBROUTINE cw_rc (oracle/cwrc_oracle.c restating kpp.f90:2152-2414) - pinned by a
numpy evaluation of the same sums, the switch table of kpp.f90:2366-2410 and golden layers (the
reference has no fixtures: "parity unpinned")."""
import os

import numpy as np

from mistra_b200 import kon
from oracle import cwrc_oracle as cwo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cwrc_layers.npz")


def inputs(n, seed):
    g = kon.kon_grid()
    d = kon.synthetic_columns(g, n, seed=seed, dry_fraction=0.4)
    r = np.random.default_rng(seed)
    ff = d["ff"] * 10.0 ** r.uniform(-1, 3, (n, 1, 1))           # from sub-threshold to cloudy water contents
    feu = np.where(r.uniform(size=n) < 0.15, r.uniform(0.2, 0.45, n), d["feu"])
    cloud = r.uniform(size=(n, 4)) < 0.5
    return g, ff, feu, cloud


def test_sums_against_numpy_and_switches():
    g, ff, feu, cloud = inputs(300, 3)
    rc, cw, cm, conv2 = cwo.cw_rc(g, ff, feu, cloud)
    nka, nkt, ka = g["nka"], g["nkt"], g["ka"]
    jt = np.arange(nkt)[None, :]
    aer = jt < np.asarray(g["kw"])[:, None]                                   # [nka,nkt]
    small = (np.arange(nka) < ka)[:, None]
    masks = [aer & small, aer & ~small, ~aer & small, ~aer & ~small]
    xpi = 4.0 / 3.0 * 3.1415926535897932
    vol = ff * xpi * g["rq"] ** 3
    for b, m in enumerate(masks):
        cwb = (vol * m).sum(axis=(1, 2))
        assert np.allclose(cw[:, b], cwb * 1e-12, rtol=1e-12)
        rcb = (vol * g["rq"] * m).sum(axis=(1, 2))
        pos = cwb > 0
        assert np.allclose(rc[pos, b], rcb[pos] / cwb[pos] * 1e-6, rtol=1e-12) and not rc[~pos, b].any()
        cmb = (ff * np.asarray(g["e"])[None, None, :] * m).sum(axis=(1, 2)) * 1e-3
        thr = 1e-1 if b < 2 else 1e2
        if b == 0:
            hum = (cloud[:, 0] & (feu >= 0.4)) | (feu >= 0.7)
        elif b == 1:
            hum = (cloud[:, 1] & (feu >= 0.42)) | (feu >= 0.75)
        else:
            hum = np.ones_like(feu, dtype=bool)
        on = (cwb >= thr) & hum & ~(feu < 0.4)
        margin = np.abs(cwb / thr - 1) > 1e-9                                # away from the threshold
        assert np.array_equal((cm[:, b] > 0)[margin], (on & (cmb > 0))[margin])
        assert np.allclose(cm[on & margin, b], cmb[on & margin], rtol=1e-12)
        assert np.allclose(conv2[on & margin, b], 1e9 / cwb[on & margin], rtol=1e-12)
        assert 0.05 < on.mean() < 0.98                                        # both states exercised
    # ifeed == 2 leaves the first dry class out
    rc2, cw2, _, _ = cwo.cw_rc(g, ff, feu, cloud, ial=2)
    m0 = masks[0].copy(); m0[0] = False
    assert np.allclose(cw2[:, 0], (vol * m0).sum(axis=(1, 2)) * 1e-12, rtol=1e-12)


def test_golden_layers():
    z = np.load(GOLD)
    g = kon.kon_grid()
    out = cwo.cw_rc(g, z["ff"], z["feu"], z["cloud"])
    for k, o in zip(("rc", "cw", "cm", "conv2"), out):
        assert np.array_equal(o, z[k])
