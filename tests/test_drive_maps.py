"""Copy lists of the *_drive routines (mistra_b200/mech/drive_maps.json extracted from aer_mk.dat, aer_km.dat
and tot.f) and their numpy restatement (oracle/drive_oracle.py): every name resolves in the mechanism, the
gather and scatter lists are each other's inverse, scatter(gather(x)) restores the mapped entries, FIX follows
aer.f:153-170 including the default-REAL literals."""
import json
import os

import numpy as np
import pytest

from mistra_b200 import drive
from mistra_b200.mechgen import mech as mechmod
from oracle import drive_oracle as dro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
J1, J5, J2, J6, NKC = 75, 20, 121, 55, 4


def name_lists(mech, seed=0):
    """Synthetic gas / radical name lists (the reference reads them from its input lists): gas-phase species of
    the mechanism plus two names it does not know (left out by match_mk_indexes)."""
    m = mechmod.load(mech)
    gasph = [s for s in m.spc_names[:m.nvar] if not (s[-2:] in ("l1", "l2", "l3", "l4"))]
    r = np.random.default_rng(seed)
    pick = list(r.permutation(gasph)[:J1 + J5 - 2])
    return pick[:J1 - 1] + ["XYZ_unknown"], pick[J1 - 1:] + ["QQ_unknown"]


def state(nl, seed):
    r = np.random.default_rng(seed)
    f = lambda *s: r.uniform(-0.2, 1.0, s)                      # some negative values to clip
    return dict(s1=f(nl, J1), s3=f(nl, J5), sl1=f(nl, NKC, J2), sion1=f(nl, NKC, J6))


def test_tables_resolve_and_invert():
    tab = json.load(open(os.path.join(ROOT, "mistra_b200", "mech", "drive_maps.json")))
    assert tab["j2"] == J2
    for mech, nbin in (("aer", 2), ("tot", 4)):
        m = mechmod.load(mech)
        g, s = tab[mech]["gather"], tab[mech]["scatter"]
        assert len(g) == 79 * nbin and sorted(map(tuple, g)) == sorted(map(tuple, s))
        assert all(e[0] in m.spc_names for e in g)
        assert len({e[0] for e in g}) == len(g)                                   # one statement per species
        assert len({(e[1], e[2], e[3]) for e in g}) == len(g)                     # and per array element
        assert all(e[0].endswith("l%d" % e[3]) for e in g)                        # bin suffix = kc
        assert max(e[2] for e in g if e[1] == "sl1") <= J2 and max(e[2] for e in g if e[1] == "sion1") <= J6


def test_unknown_species_names_abort_like_the_reference():
    """match_mk_indexes stops the model on a user species the mechanism does not know (utils.f90:1052-1056)."""
    gn, rn = name_lists("aer")
    with pytest.raises(ValueError, match="XYZ_unknown"):
        drive.drive_map("aer", gn, rn)
    assert len(drive.drive_map("aer", gn, rn, allow_unmatched=True)["kpp"]) == len(gn) + len(rn) - 2 + 79 * 2


def test_oracle_gather_scatter():
    for mech in ("aer", "tot", "gas"):
        gn, rn = name_lists(mech)
        mp = drive.drive_map(mech, gn, rn, allow_unmatched=True)
        m = mechmod.load(mech)
        assert mp["indf_o2"] and mp["indf_h2o"] and mp["indf_n2"]
        assert [bool(x) for x in mp["indf_h2ol"]] == {"aer": [1, 1, 0, 0], "tot": [1, 1, 1, 1], "gas": [0, 0, 0, 0]}[mech]
        nl = 40
        st = state(nl, 3)
        layer = np.random.default_rng(4).permutation(nl)[:25].astype(np.int64)
        r = np.random.default_rng(5)
        air, h2o = r.uniform(30, 45, 25), r.uniform(0.1, 0.6, 25)
        cvv = np.where(r.uniform(size=(25, 4)) < 0.3, 0.0, r.uniform(1e3, 1e6, (25, 4)))
        var0, fix0 = np.full((25, m.nvar), -7.0), np.full((25, m.nfix), -7.0)
        var, fix, sl1c, sion1c = dro.gather(mp, layer, st["s1"], st["s3"], st["sl1"], st["sion1"], air, h2o, cvv, var0, fix0)
        # mapped entries carry the (clamped) model values, the others keep the caller's
        mapped = np.zeros(m.nvar + m.nfix, dtype=bool); mapped[mp["kpp"] - 1] = True
        assert np.all(var[:, ~mapped[:m.nvar]] == -7.0) and (~mapped[:m.nvar]).any()
        assert np.array_equal(sl1c[layer], np.maximum(0, st["sl1"][layer])) and (sl1c >= 0)[layer].all()
        untouched = np.setdiff1d(np.arange(nl), layer)
        assert np.array_equal(sl1c[untouched], st["sl1"][untouched])
        assert np.array_equal(fix[:, mp["indf_o2"] - 1], float(np.float32(0.21)) * air)
        if mech != "gas":
            p = mp["indf_h2ol"][0] - 1
            assert np.array_equal(fix[:, p] == 0, cvv[:, 0] == 0)
            on = cvv[:, 0] > 0
            assert np.array_equal(fix[on, p], float(np.float32(55.55)) / cvv[on, 0])
        # scatter of the gathered vector restores every mapped element (clipped), whatever the arrays held before
        junk = state(nl, 9)
        s1, s3, sl1, sion1 = dro.scatter(mp, layer, junk["s1"], junk["s3"], junk["sl1"], junk["sion1"], var, fix)
        rows_new = [s1, s3, sl1.reshape(nl, -1), sion1.reshape(nl, -1)]
        rows_old = [st["s1"], st["s3"], sl1c.reshape(nl, -1), sion1c.reshape(nl, -1)]
        for kp, ar, of in zip(mp["kpp"], mp["arr"], mp["off"]):
            assert np.array_equal(rows_new[ar][layer, of], np.maximum(0.0, rows_old[ar][layer, of]))
        assert all((x[layer] >= 0).all() for x in rows_new)
