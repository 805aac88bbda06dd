"""CPU oracle of kpp_driver's layer loop (oracle/driver_oracle.py restating kpp.f90:4305-4470) against the records of the
reference's own statements (tests/golden/driver_reference.npz, made by tests/golden/make_driver_reference.py: which
*_drive the reference called for every layer, with which arguments)."""
import os

import numpy as np

from oracle import driver_oracle as do

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "driver_reference.npz")


def case(z, tag):
    g = lambda k: z["%s_%s" % (tag, k)]
    fl = g("flags")
    cfg = dict(nf=int(g("sizes")[2]), halo=bool(fl[0]), iod=bool(fl[1]), lpBuys13_0D=bool(fl[2]), neula=int(fl[3]),
               box=bool(fl[4]), n_bl=int(fl[5]), kinv=int(fl[6]), dt_ch=float(g("dt_ch")))
    adv_row = np.array([int(g("ind_gas_rev")[s]) - 1 if s else -1 for s in g("nindadv")], dtype=np.int32)
    ins = {k: g("in_" + k) for k in ("u0", "t", "p", "rho", "cm3", "am3", "xm1", "conv2", "cm", "cloud", "photol_j", "s1", "s3")}
    return cfg, adv_row, g("xadv"), ins, g


def check_case(fn, z, tag):
    cfg, adv_row, xadv, ins, g = case(z, tag)
    o = fn(cfg, ins["u0"], ins["t"], ins["p"], ins["rho"], ins["cm3"], ins["am3"], ins["xm1"], ins["conv2"], ins["cm"],
           ins["cloud"], ins["photol_j"], adv_row, xadv, ins["s1"], ins["s3"])
    n = ins["t"].shape[1]
    L = g("rec_col") * n + g("rec_k") - 1
    assert np.array_equal(np.sort(L), np.nonzero(o["mech"] >= 0)[0])              # the same layers are integrated ...
    assert np.array_equal(o["mech"][L], g("rec_mech"))                            # ... by the same mechanism
    for m in range(3):
        assert np.array_equal(o["layers"][m], np.sort(L[g("rec_mech") == m]))
    assert np.array_equal(o["cb1"][L], g("rec_cb1")) and np.array_equal(o["air"][L], g("rec_air"))
    assert np.array_equal(o["h2o"][L], g("rec_h2o")) and np.array_equal(o["ph_rat"][L], g("rec_ph"))
    assert np.array_equal(o["scal"][L, 1:], g("rec_scal")) and np.array_equal(o["cvv"][L], g("rec_scal")[:, 8:])
    assert np.all(o["scal"][L, 0] == 6.022140857e+23 * 1.e-6)
    assert np.array_equal(o["cloud"], g("out_cloud")) and np.array_equal(o["s1"], g("out_s1")) and np.array_equal(o["s3"], g("out_s3"))
    assert np.all(g("rec_t") == np.array([0.0, cfg["dt_ch"]]))
    return o


def test_oracle_reproduces_the_reference_statements():
    z = np.load(GOLD)
    o = check_case(do.layers, z, "a")
    assert all(len(x) > 0 for x in o["layers"]) and (o["ph_rat"] != 0).any()
    dark = z["a_in_u0"] < 3.48e-2
    n = z["a_in_t"].shape[1]
    assert dark.any() and not o["ph_rat"].reshape(len(dark), n, -1)[dark].any()   # night columns: no photolysis
    assert (z["a_out_s1"] != np.where(z["a_in_s1"] < 0, 0, z["a_in_s1"])).any()   # the advection source acted
    check_case(do.layers, z, "b")                                                 # halo off, Buys13 threshold, no advection
    o = check_case(do.layers, z, "c")                                             # box run: one level per column
    assert (o["mech"] >= 0).sum() == z["c_in_t"].shape[0]
