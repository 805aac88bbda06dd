"""CUDA condensation step (subkon + advec) vs the CPU oracle through the C ABI (-m gpu).
Contract (include/mistra_kon.h): same statement order in binary64 without FMA; exp/pow
come from the CUDA math library (<= 2 ulp from libm) and dwsum is summed per dry class
first, so: identical iteration counts, to/xm1o to 1e-12 relative, the spectrum to 1e-9
relative to the layer's largest bin content."""
import os

import numpy as np
import pytest

from mistra_b200 import kon
from oracle import kon_oracle as ko

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kon_layers.npz")
KEYS = ("ffk", "totr", "dfdt", "feualt", "pp", "to", "tn", "xm1o", "xm1n", "kr")


def check(out, ref, min_same=0.99):
    ffk, to, xm1o, st = out
    ffk_r, to_r, xm1o_r, st_r = ref
    same = st == st_r
    assert same.mean() >= min_same, "iteration counts differ in %.1f%% of the layers" % (100 * (1 - same.mean()))
    assert np.allclose(to[same], to_r[same], rtol=1e-12) and np.allclose(xm1o[same], xm1o_r[same], rtol=1e-11)
    scale = np.abs(ffk_r).max(axis=(1, 2), keepdims=True) + 1e-300
    assert (np.abs(ffk - ffk_r) / scale)[same].max() <= 1e-9
    # layers whose secant iteration stopped one step earlier/later: converged to the same answer
    assert np.allclose(to[~same], to_r[~same], rtol=1e-7)


def test_golden_layers(cuda_device, kpp):
    g = np.load(GOLD)
    grid = kon.kon_grid()
    out = kon.subkon(grid, float(g["dt"]), *[g[k] for k in KEYS])
    check(out, (g["ffk_out"], g["to_out"], g["xm1o_out"], g["status"]), min_same=1.0)


@pytest.mark.parametrize("seed,dt", [(1, 10.0), (2, 10.0), (3, 2.0), (4, 60.0)])
def test_synthetic_layers_vs_oracle(cuda_device, kpp, seed, dt):
    grid = kon.kon_grid()
    d = kon.synthetic_layers(grid, 400, seed=seed)
    ref = ko.subkon(grid, dt, *[d[k] for k in KEYS])
    out = kon.subkon(grid, dt, *[d[k] for k in KEYS])
    check(out, ref)
    assert (out[3] >= 1).mean() > 0.95


def test_other_grids(cuda_device, kpp):
    for args in ((0.01, 2.0, 0.01, 80.0), (0.01, 10.0, 0.01, 550.0)):        # BTZ96 / Bott2020 namelists
        grid = kon.kon_grid(*args)
        d = kon.synthetic_layers(grid, 96, seed=5)
        check(kon.subkon(grid, 10.0, *[d[k] for k in KEYS]), ko.subkon(grid, 10.0, *[d[k] for k in KEYS]))


def test_edge_cases(cuda_device, kpp):
    grid = kon.kon_grid()
    d = kon.synthetic_layers(grid, 6, seed=6)
    e = {k: v[:0] for k, v in d.items()}
    out = kon.subkon(grid, 10.0, *[e[k] for k in KEYS])                       # empty batch
    assert out[0].shape[0] == 0
    d["ffk"][1] = 0.0                                                         # no particles: nothing happens
    d["feualt"][2] = 0.75; d["xm1o"][2] *= 0.6; d["xm1n"][2] *= 0.6           # very dry air: strong evaporation
    d["feualt"][3] = 1.05; d["xm1o"][3] *= 1.05; d["xm1n"][3] *= 1.05         # 5 % supersaturation
    ref = ko.subkon(grid, 10.0, *[d[k] for k in KEYS])
    out = kon.subkon(grid, 10.0, *[d[k] for k in KEYS])
    assert np.array_equal(out[3], ref[3])
    check(out, ref, min_same=1.0)
    assert not out[0][1].any() and out[1][1] == d["tn"][1]
    # a non-finite layer is reported like the oracle does and leaves its neighbours alone
    d["to"][4] = np.nan
    ref = ko.subkon(grid, 10.0, *[d[k] for k in KEYS])
    out = kon.subkon(grid, 10.0, *[d[k] for k in KEYS])
    ok = np.arange(6) != 4
    assert np.array_equal(out[3][ok], ref[3][ok]) and out[3][4] == ref[3][4]
    assert np.allclose(out[1][ok], ref[1][ok], rtol=1e-12)
    with pytest.raises(kpp.KppError):
        kon.subkon(grid, -1.0, *[d[k] for k in KEYS])


def test_large_batch_permutation_and_device_entry(cuda_device, kpp):
    import torch
    grid = kon.kon_grid()
    d = kon.synthetic_layers(grid, 1500, seed=7)                              # > resident CTAs
    out = kon.subkon(grid, 10.0, *[d[k] for k in KEYS])
    p = np.random.default_rng(0).permutation(1500)
    out2 = kon.subkon(grid, 10.0, *[d[k][p] for k in KEYS])
    for a, b in zip(out, out2):
        assert np.array_equal(a[p], b)
    # budgets at full size (no oracle needed)
    assert np.allclose(out[0].sum(axis=2), d["ffk"].sum(axis=2), rtol=1e-12, atol=1e-300)
    dw = ((out[0] - d["ffk"]) * grid["e"][None, None, :]).sum(axis=(1, 2))
    rho = d["pp"] / (8.3144743 / 28.96546e-3 * d["to"] * (1.0 + 0.61 * d["xm1o"]))
    conv = out[3] >= 1
    assert np.allclose((d["xm1n"] - out[2])[conv], (dw / rho)[conv], rtol=1e-9, atol=1e-18)
    # device entry
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(cuda_device) for k in KEYS}
    st = torch.zeros(1500, dtype=torch.int32, device=cuda_device)
    n0 = kon.launch_count()
    kon.subkon_device(grid, 10.0, *[t[k] for k in KEYS], status=st)
    torch.cuda.synchronize()
    assert kon.launch_count() == n0 + 1
    assert np.array_equal(t["ffk"].cpu().numpy(), out[0]) and np.array_equal(st.cpu().numpy(), out[3])
    assert np.array_equal(t["to"].cpu().numpy(), out[1])


LKEYS = ("ff", "t", "talt", "xm1", "xm1a", "feu", "dfddt", "xm2", "dtcon")
SKEYS = ("vol1_a", "vol1_d", "part_o_a", "part_o_d", "part_n_a", "part_n_d", "vol2", "pntot")


def check_layers(o, r, chem):
    dry = r["status"] == 0
    same = o["status"] == r["status"]
    assert same.mean() >= 0.99 and np.array_equal(o["status"][dry], r["status"][dry])
    # dry branch: Newton iteration with the CUDA log/exp - same bins, same numbers
    assert np.array_equal(o["ff"][dry], r["ff"][dry])
    for k in ("t", "talt", "xm1", "xm1a", "dtcon"):
        assert np.array_equal(o[k][dry], r[k][dry])
    assert np.allclose(o["feu"][dry], r["feu"][dry], rtol=1e-14) and np.allclose(o["xm2"][dry], r["xm2"][dry], rtol=1e-13)
    hs = same & ~dry
    for k in ("t", "talt", "xm1", "xm1a", "feu", "xm2"):
        assert np.allclose(o[k][hs], r[k][hs], rtol=1e-11), k
    assert np.allclose(o["dfddt"][hs], r["dfddt"][hs], rtol=1e-6, atol=1e-12)
    assert np.allclose(o["dtcon"][hs], r["dtcon"][hs], rtol=1e-6, atol=1e-12)
    scale = np.abs(r["ff"]).max(axis=(1, 2), keepdims=True) + 1e-300
    assert (np.abs(o["ff"] - r["ff"]) / scale)[hs].max() <= 1e-9
    if chem:
        for k in SKEYS:
            assert np.allclose(o[k][same], r[k][same], rtol=1e-9, atol=1e-12 * np.abs(r[k]).max()), k


@pytest.mark.parametrize("chem", [True, False])
def test_kon_layer_loop_vs_oracle(cuda_device, kpp, chem):
    grid = kon.kon_grid()
    st = kon.synthetic_columns(grid, 500, seed=21, dry_fraction=0.35)
    r = ko.layers(grid, 10.0, chem, st)
    o = kon.layers(grid, 10.0, chem, st)
    check_layers(o, r, chem)
    assert (r["status"] == 0).sum() > 100 and (r["status"] >= 1).sum() > 200
    if not chem:
        assert not o["vol2"].any()


def test_kon_layer_loop_edges(cuda_device, kpp):
    grid = kon.kon_grid(0.01, 2.0, 0.01, 80.0)                                # BTZ96 grid
    st = kon.synthetic_columns(grid, 40, seed=22, dry_fraction=0.5)
    st["ff"][3] = 0.0                                                         # empty layers, dry and humid
    st["feu"][4] = 0.6999999; st["feu"][5] = 0.7                             # either side of the branch
    r = ko.layers(grid, 10.0, True, st)
    o = kon.layers(grid, 10.0, True, st)
    check_layers(o, r, True)
    assert r["status"][4] == 0 and r["status"][5] != 0
    e = {k: v[:0] for k, v in st.items()}
    assert kon.layers(grid, 10.0, True, e)["ff"].shape[0] == 0
    bad = dict(grid); bad["kw"] = grid["kw"] + 1000
    with pytest.raises(kpp.KppError):
        kon.layers(bad, 10.0, True, st)
