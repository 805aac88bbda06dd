"""SUBROUTINE cw_rc on the device (row N2, first piece; include/mistra_cwrc.h) vs the CPU oracle
(-m gpu): sums to 1e-13 relative (per-class partial sums, documented), switches identical."""
import os

import numpy as np
import pytest

from mistra_b200 import cwrc, kon
from oracle import cwrc_oracle as cwo
from tests.test_cwrc_oracle import inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cwrc_layers.npz")


def compare(out, ref):
    for name, o, r in zip(("rc", "cw", "cm", "conv2"), out, ref):
        assert np.array_equal(o == 0, r == 0), name                  # the same bins are switched off
        assert np.allclose(o, r, rtol=1e-13, atol=0), name


def test_golden_layers(cuda_device, kpp):
    z = np.load(GOLD)
    g = kon.kon_grid()
    n0 = cwrc.launch_count()
    out = cwrc.cw_rc(g, z["ff"], z["feu"], z["cloud"])
    assert cwrc.launch_count() == n0 + 1
    compare(out, tuple(z[k] for k in ("rc", "cw", "cm", "conv2")))


@pytest.mark.parametrize("n,seed,ial", [(2000, 5, 1), (333, 6, 2), (1, 7, 1)])
def test_synthetic_layers_vs_oracle(cuda_device, kpp, n, seed, ial):
    g, ff, feu, cloud = inputs(n, seed)
    compare(cwrc.cw_rc(g, ff, feu, cloud, ial=ial), cwo.cw_rc(g, ff, feu, cloud, ial=ial))


def test_other_grid_device_entry_and_edges(cuda_device, kpp):
    import torch
    g = kon.kon_grid(0.01, 2.0, 0.01, 80.0)                          # BTZ96 grid
    n = 5000
    d = kon.synthetic_columns(g, n, seed=9, dry_fraction=0.3)
    ff = d["ff"] * 50.0
    ff[3] = 0.0                                                      # empty layer: rc = cw = 0, switched off
    feu = d["feu"]
    cloud = np.zeros((n, 4), dtype=np.int32); cloud[::2] = 1
    ref = cwo.cw_rc(g, ff, feu, cloud)
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)
    gd = {"nka": g["nka"], "nkt": g["nkt"], "ka": g["ka"], "kw": t(g["kw"], np.int32), "e": t(g["e"]), "rq": t(g["rq"])}
    outs = [torch.full((n, 4), -1.0, dtype=torch.float64, device=cuda_device) for _ in range(4)]
    cwrc.cw_rc_device(gd, t(ff), t(feu), t(cloud, np.int32), *outs)
    torch.cuda.synchronize()
    compare(tuple(o.cpu().numpy() for o in outs), ref)
    assert not ref[1][3].any() and not outs[1][3].any().item()
    e = cwrc.cw_rc(g, ff[:0], feu[:0], cloud[:0])
    assert e[0].shape == (0, 4)
    with pytest.raises(ValueError):
        cwrc.cw_rc(g, ff[:, :10], feu, cloud)
    with pytest.raises(Exception):
        cwrc.cw_rc(g, ff, feu, cloud, ial=3)
