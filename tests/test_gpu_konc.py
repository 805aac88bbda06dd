"""SUBROUTINE konc on the device (row N3, include/mistra_konc.h) vs the CPU oracle (-m gpu):
bit-exact (binary64, same statement order, no FMA contraction), warning counts identical."""
import os

import numpy as np
import pytest

from mistra_b200 import konc
from oracle import konc_oracle as kco

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "konc_layers.npz")


def test_golden_layers(cuda_device, kpp):
    g = np.load(GOLD)
    sums = {k: g[k] for k in konc.SUMS}
    n0 = konc.launch_count()
    sl1, sion1, warn = konc.konc(int(g["ka"]), sums, g["vol2"], g["pntot"], g["sl1"], g["sion1"])
    assert konc.launch_count() == n0 + 1
    assert np.array_equal(sl1, g["sl1_out"]) and np.array_equal(sion1, g["sion1_out"])
    assert np.array_equal(warn, g["warn"])


@pytest.mark.parametrize("n,nka,ka,j2,j6,seed", [(3000, 70, 32, 121, 55, 1), (257, 40, 0, 17, 3, 2),
                                                 (100, 96, 96, 1, 200, 3), (1, 70, 32, 121, 55, 4)])
def test_synthetic_layers_vs_oracle(cuda_device, kpp, n, nka, ka, j2, j6, seed):
    d = konc.synthetic_sums(n, nka=nka, ka=ka, seed=seed, j2=j2, j6=j6)
    ref = kco.konc(d["ka"], d["sums"], d["vol2"], d["pntot"], d["sl1"], d["sion1"])
    out = konc.konc(d["ka"], d["sums"], d["vol2"], d["pntot"], d["sl1"], d["sion1"])
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)


def test_device_entry_large_batch_and_edges(cuda_device, kpp):
    import torch
    n = 40000                                                 # 0.6 GB of sums + species, more layers than CTAs
    d = konc.synthetic_sums(n, seed=11)
    ref = kco.konc(d["ka"], d["sums"], d["vol2"], d["pntot"], d["sl1"], d["sion1"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    sums = {k: t(v) for k, v in d["sums"].items()}
    sl1, sion1 = t(d["sl1"]), t(d["sion1"])
    warn = torch.zeros((n, 3), dtype=torch.int32, device=cuda_device)
    konc.konc_device(d["ka"], sums, t(d["vol2"]), t(d["pntot"]), sl1, sion1, warn=warn)
    torch.cuda.synchronize()
    assert np.array_equal(sl1.cpu().numpy(), ref[0]) and np.array_equal(sion1.cpu().numpy(), ref[1])
    assert np.array_equal(warn.cpu().numpy(), ref[2])
    # size-independent property at full size: with enough droplets and clean sums, each bin pair keeps its total
    # empty batch, bad sizes
    e = konc.synthetic_sums(0)
    out = konc.konc(e["ka"], e["sums"], e["vol2"], e["pntot"], e["sl1"], e["sion1"])
    assert out[0].shape == (0, 4, 121)
    with pytest.raises(Exception):
        konc.konc(71, d["sums"], d["vol2"], d["pntot"], d["sl1"], d["sion1"])     # ka > nka
    with pytest.raises(ValueError):
        konc.konc(32, d["sums"], d["vol2"][:, :3], d["pntot"], d["sl1"], d["sion1"])


def test_chain_kon_layers_to_konc(cuda_device, kpp):
    """The bin sums mistra_kon_layers(chem) leaves behind drive konc: CUDA chain vs the oracle of
    konc on the same sums (bit-exact), and the sums themselves against the oracle of kon."""
    from mistra_b200 import kon
    from oracle import kon_oracle as kno
    g = kon.kon_grid()
    st = kon.synthetic_columns(g, 64, seed=3, dry_fraction=0.3)
    o = kon.layers(g, 10.0, True, st)
    r = kno.layers(g, 10.0, True, st)
    rng = np.random.default_rng(8)
    sl1 = 10.0 ** rng.uniform(-12, -6, (64, 4, 121))
    sion1 = 10.0 ** rng.uniform(-12, -6, (64, 4, 55))
    sums = {k: o[k] for k in konc.SUMS}
    out = konc.konc(g["ka"], sums, o["vol2"], o["pntot"], sl1, sion1)
    ref = kco.konc(g["ka"], sums, o["vol2"], o["pntot"], sl1, sion1)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)
    moved = np.abs(out[0] - sl1).max(axis=(1, 2)) > 0
    assert moved.sum() > 10                                    # condensation did move particles across kw
    for k in konc.SUMS:                                        # kon's sums: rounding-level agreement with its oracle
        assert np.allclose(o[k], r[k], rtol=1e-9, atol=1e-12 * np.abs(r[k]).max())
