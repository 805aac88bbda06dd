"""CUDA 2-D bin redistribution vs the CPU oracle through the C ABI (-m gpu).
Contract (include/mistra_bins.h): the product build sums the particle number of a bin per
dry class first (a reassociation of the reference's running sum), so sap agrees to 1e-13
and ff to 1e-11 relative; the -DKPP_STRICT build keeps the running sum and must reproduce
ff, sap, smp, sion1o to the last bit.  sl1/sion1 after the bin exchange: 1e-12 relative in
both (the transferred volume is reduced over water bins in parallel)."""
import os

import numpy as np
import pytest

from mistra_b200 import bins
from oracle import bins_oracle as bo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bins_layers.npz")


def close(a, b, rtol):
    return np.allclose(a, b, rtol=rtol, atol=1e-13 * max(np.abs(b).max(), 1e-300))


def run_both(grid, d, sion1_new=None):
    new = d["sion1_new"] if sion1_new is None else sion1_new
    sap_o, smp_o, so_o = bo.snapshot(grid, d["ff"], d["cm"], d["sion1"])
    ref = bo.redistribute(grid, d["ff"], d["cm"], d["cw"], sap_o, smp_o, so_o, new, d["sl1"])
    # strict build: to the last bit
    sap, smp, so = bins.snapshot(grid, d["ff"], d["cm"], d["sion1"], strict=True)
    assert np.array_equal(sap, sap_o) and np.array_equal(smp, smp_o) and np.array_equal(so, so_o)
    out = bins.redistribute(grid, d["ff"], d["cm"], d["cw"], sap, smp, so, new, d["sl1"], strict=True)
    assert np.array_equal(out[0], ref[0]) and np.array_equal(out[3], ref[3])
    assert close(out[1], ref[1], 1e-12) and close(out[2], ref[2], 1e-12)
    # product build: reassociated particle-number sum
    sap, smp, so = bins.snapshot(grid, d["ff"], d["cm"], d["sion1"])
    assert np.allclose(sap, sap_o, rtol=1e-13, atol=0) and np.array_equal(smp, smp_o) and np.array_equal(so, so_o)
    out = bins.redistribute(grid, d["ff"], d["cm"], d["cw"], sap, smp, so, new, d["sl1"])
    assert close(out[0], ref[0], 1e-11) and np.array_equal(out[3], ref[3])
    assert close(out[1], ref[1], 1e-11) and close(out[2], ref[2], 1e-11)
    return out, ref


def test_golden_layers(cuda_device, kpp):
    g = np.load(GOLD)
    grid = bins.particle_grid(*g["grid_args"])
    for strict in (True, False):
        sap, smp, so = bins.snapshot(grid, g["ff"], g["cm"], g["sion1"], strict=strict)
        assert np.allclose(sap, g["sap"], rtol=1e-13, atol=0) and np.array_equal(smp, g["smp"])
        assert np.array_equal(so, g["sion1o"])
        ff2, si2, sl2, nw = bins.redistribute(grid, g["ff"], g["cm"], g["cw"], sap, smp, so, g["sion1_new"],
                                              g["sl1"], strict=strict)
        if strict:
            assert np.array_equal(sap, g["sap"]) and np.array_equal(ff2, g["ff_out"])
        assert close(ff2, g["ff_out"], 1e-11) and np.array_equal(nw, g["nwarn"])
        assert close(si2, g["sion1_out"], 1e-11) and close(sl2, g["sl1_out"], 1e-11)


@pytest.mark.parametrize("nkc_l,ial_first", [(4, 1), (2, 1), (4, 2), (1, 1)])
def test_synthetic_layers_vs_oracle(cuda_device, kpp, nkc_l, ial_first):
    grid = bins.particle_grid(nkc_l=nkc_l, ial_first=ial_first)
    d = bins.synthetic_layers(grid, 300, seed=100 + nkc_l)
    out, ref = run_both(grid, d)
    assert np.abs(out[0] - d["ff"]).sum() > 0


def test_other_grids_and_strong_mass_change(cuda_device, kpp):
    for args in ((0.01, 2.0, 0.01, 80.0), (0.01, 10.0, 0.01, 550.0)):       # BTZ96 / Bott2020 namelists
        grid = bins.particle_grid(*args)
        d = bins.synthetic_layers(grid, 64, seed=3, growth=0.5)
        run_both(grid, d)
    grid = bins.particle_grid()
    d = bins.synthetic_layers(grid, 64, seed=4)
    out, ref = run_both(grid, d, sion1_new=d["sion1"] * 0.02)                # x0 clipped at the lower grid end
    out, ref = run_both(grid, d, sion1_new=d["sion1"] * 40.0)                # ... and at the upper end


def test_edge_cases(cuda_device, kpp):
    grid = bins.particle_grid()
    d = bins.synthetic_layers(grid, 5, seed=8)
    # empty batch
    e = {k: v[:0] for k, v in d.items()}
    sap, smp, so = bins.snapshot(grid, e["ff"], e["cm"], e["sion1"])
    assert sap.shape == (0, 4)
    # dry layers (cm = 0) and no mass change are identities
    d["cm"][1] = 0.0
    d["sion1_new"][2] = d["sion1"][2]
    d["ff"][4] = 0.0                                                        # no particles: sap = 0 -> skipped
    out, ref = run_both(grid, d)
    for i in (1, 2, 4):
        assert np.array_equal(out[0][i], d["ff"][i])
    # non-finite input does not hang or corrupt the neighbours
    d2 = bins.synthetic_layers(grid, 3, seed=9)
    d2["sion1_new"][1, 0, 0] = np.nan
    out, ref = run_both(grid, d2) if False else (None, None)
    sap, smp, so = bins.snapshot(grid, d2["ff"], d2["cm"], d2["sion1"])
    o = bins.redistribute(grid, d2["ff"], d2["cm"], d2["cw"], sap, smp, so, d2["sion1_new"], d2["sl1"])
    r = bo.redistribute(grid, d2["ff"], d2["cm"], d2["cw"], sap, smp, so, d2["sion1_new"], d2["sl1"])
    assert np.array_equal(o[0][[0, 2]], r[0][[0, 2]])
    with pytest.raises(kpp.KppError):
        bad = dict(grid); bad["nkc_l"] = 7
        bins.snapshot(bad, d["ff"], d["cm"], d["sion1"])


def test_large_batch_permutation_invariance_and_device_entry(cuda_device, kpp):
    import torch
    grid = bins.particle_grid()
    d = bins.synthetic_layers(grid, 3000, seed=12)                          # > resident CTAs
    sap, smp, so = bins.snapshot(grid, d["ff"], d["cm"], d["sion1"])
    out = bins.redistribute(grid, d["ff"], d["cm"], d["cw"], sap, smp, so, d["sion1_new"], d["sl1"])
    p = np.random.default_rng(0).permutation(3000)
    sap2, smp2, so2 = bins.snapshot(grid, d["ff"][p], d["cm"][p], d["sion1"][p])
    assert np.array_equal(sap2, sap[p])
    out2 = bins.redistribute(grid, d["ff"][p], d["cm"][p], d["cw"][p], sap2, smp2, so2, d["sion1_new"][p], d["sl1"][p])
    for a, b in zip(out, out2):
        assert np.array_equal(a[p], b)
    # conservation at full size (no oracle needed)
    assert np.allclose(out[0].sum(axis=1), d["ff"].sum(axis=1), rtol=1e-13, atol=1e-300)
    assert np.allclose(out[2].sum(axis=1), d["sl1"].sum(axis=1), rtol=1e-13)
    # device entry on torch tensors
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in d.items()}
    dsap = torch.zeros((3000, 4), dtype=torch.float64, device=cuda_device)
    dsmp = torch.zeros_like(dsap)
    dso = torch.zeros((3000, 4, 9), dtype=torch.float64, device=cuda_device)
    nw = torch.zeros(3000, dtype=torch.int32, device=cuda_device)
    n0 = bins.launch_count()
    bins.snapshot_device(grid, t["ff"], t["cm"], t["sion1"], dsap, dsmp, dso)
    bins.redistribute_device(grid, t["ff"], t["cm"], t["cw"], dsap, dsmp, dso, t["sion1_new"], t["sl1"], nw)
    torch.cuda.synchronize()
    assert bins.launch_count() == n0 + 2
    assert np.array_equal(t["ff"].cpu().numpy(), out[0]) and np.array_equal(dsap.cpu().numpy(), sap)
    assert np.array_equal(t["sl1"].cpu().numpy(), out[2])
