"""SUBROUTINE fast_k_mt_a / fast_k_mt_t on the device (row N2, second piece; include/mistra_fastkmt.h)
vs the CPU oracle (-m gpu): integrals to 1e-13 relative (per-warp partial sums, documented), the same
entries written, everything else untouched bit for bit."""
import os

import numpy as np
import pytest

from mistra_b200 import fastkmt, kon
from oracle import fastkmt_oracle as fko
from tests.test_fastkmt_oracle import call, inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fastkmt_layers.npz")


def compare(out, ref, prev):
    for name, o, r, p0 in zip(("xkmt", "vt"), out, ref, prev):
        same = r == p0                                   # entries the reference did not assign
        assert np.array_equal(o[same], p0[same]), name
        assert np.allclose(o, r, rtol=1e-13, atol=0), name


def test_golden_layers(cuda_device, kpp):
    z = np.load(GOLD)
    x = {k: z[k] for k in ("lex", "ff", "freep", "t", "p", "cw", "cm", "alpha", "vmean", "xkmt", "vt")}
    x["g"] = kon.kon_grid()
    n0 = fastkmt.launch_count()
    out = call(fastkmt.fast_k_mt, x)
    assert fastkmt.launch_count() == n0 + 1
    compare(out, (z["xkmt_out"], z["vt_out"]), (z["xkmt"], z["vt"]))


@pytest.mark.parametrize("n,seed,mech,nkc_l,ial,dense", [(700, 5, "aer", 4, 1, 0.5), (333, 6, "aer", 2, 2, 0.05),
                                                         (400, 8, "tot", 4, 1, 1.0), (1, 7, "tot", 4, 1, 0.5)])
def test_synthetic_layers_vs_oracle(cuda_device, kpp, n, seed, mech, nkc_l, ial, dense):
    x = inputs(n, seed, mech, dense=dense)
    ref = call(fko.fast_k_mt, x, nkc_l=nkc_l, ial=ial)
    assert (ref[0] != x["xkmt"]).any() and (ref[1] != x["vt"]).any()
    compare(call(fastkmt.fast_k_mt, x, nkc_l=nkc_l, ial=ial), ref, (x["xkmt"], x["vt"]))


def test_other_grid_device_entry_and_edges(cuda_device, kpp):
    import torch
    g = kon.kon_grid(0.01, 2.0, 0.01, 80.0)                          # BTZ96 grid
    n = 1500
    x = inputs(n, 9, "aer", g=g)
    x["ff"][3] = 0.0                                                 # empty layer: sums are zero
    x["cm"][5] = 0.0                                                 # no chemistry anywhere: only vt
    x["cw"][7] = 0.0                                                 # nothing assigned
    ref = call(fko.fast_k_mt, x)
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)
    gd = {"nka": g["nka"], "nkt": g["nkt"], "ka": g["ka"], "kw": t(g["kw"], np.int32), "rq": t(g["rq"])}
    xk, vt = t(x["xkmt"]), t(x["vt"])
    fastkmt.fast_k_mt_device(gd, t(x["lex"], np.int32), t(x["ff"]), t(x["freep"]), t(x["t"]), t(x["p"]), t(x["cw"]),
                             t(x["cm"]), t(x["alpha"]), t(x["vmean"]), xk, vt)
    torch.cuda.synchronize()
    out = (xk.cpu().numpy(), vt.cpu().numpy())
    compare(out, ref, (x["xkmt"], x["vt"]))
    assert np.array_equal(out[0][5], x["xkmt"][5]) and np.array_equal(out[0][7], x["xkmt"][7])
    assert np.array_equal(out[1][7], x["vt"][7]) and (out[1][5] != x["vt"][5]).any()
    e = fastkmt.fast_k_mt(g, x["lex"], x["ff"][:0], x["freep"][:0], x["t"][:0], x["p"][:0], x["cw"][:0], x["cm"][:0],
                          x["alpha"][:0], x["vmean"][:0], x["xkmt"][:0], x["vt"][:0])
    assert e[0].shape == (0, 4, 262)
    with pytest.raises(ValueError):
        fastkmt.fast_k_mt(g, x["lex"], x["ff"][:, :10], x["freep"], x["t"], x["p"], x["cw"], x["cm"], x["alpha"],
                          x["vmean"], x["xkmt"], x["vt"])
    with pytest.raises(ValueError):
        fastkmt.fast_k_mt(g, np.array([0, 5]), x["ff"], x["freep"], x["t"], x["p"], x["cw"], x["cm"], x["alpha"],
                          x["vmean"], x["xkmt"], x["vt"])
    with pytest.raises(Exception):
        call(fastkmt.fast_k_mt, x, ial=3)
