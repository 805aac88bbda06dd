"""The layer loop of kpp_driver on the device (row a15; include/mistra_driver.h) vs the records of the reference's own
statements and vs the CPU oracle (-m gpu): bit for bit."""
import numpy as np
import pytest

from mistra_b200 import driver as dm
from oracle import driver_oracle as do
from tests.test_driver_oracle import GOLD, check_case

pytestmark = pytest.mark.gpu
KEYS = ("cb1", "scal", "ph_rat", "air", "h2o", "cvv", "mech", "cloud", "s1", "s3")


def test_reference_statement_records(cuda_device, kpp):
    z = np.load(GOLD)
    n0 = dm.launch_count()
    check_case(dm.layers, z, "a")
    assert dm.launch_count() == n0 + 6                  # clip s1, clip s3, layers, count, scan, write
    check_case(dm.layers, z, "b")
    check_case(dm.layers, z, "c")


def inputs(ncol, seed, n=150, nf=100):
    d = dm.synthetic_columns(ncol, n, nf, 4, 93, 24, seed)
    d["u0"] = np.random.default_rng(seed).uniform(-0.2, 1.0, ncol)
    adv_row = np.array([4, -1, 17, 4, 60], dtype=np.int32)
    xadv = np.array([1e-9, 2e-9, -3e-10, 5e-10, 1e-8])
    cfg = dict(nf=nf, halo=True, iod=True, lpBuys13_0D=False, neula=0, box=False, n_bl=2, kinv=70, dt_ch=10.0)
    return cfg, d, adv_row, xadv


def call(fn, cfg, d, adv_row, xadv):
    return fn(cfg, d["u0"], d["t"], d["p"], d["rho"], d["cm3"], d["am3"], d["xm1"], d["conv2"], d["cm"], d["cloud"],
              d["photol_j"], adv_row, xadv, d["s1"], d["s3"])


@pytest.mark.parametrize("ncol,seed", [(40, 3), (1, 4), (333, 5)])
def test_columns_vs_oracle(cuda_device, kpp, ncol, seed):
    cfg, d, adv_row, xadv = inputs(ncol, seed)
    o, r = call(dm.layers, cfg, d, adv_row, xadv), call(do.layers, cfg, d, adv_row, xadv)
    for k in KEYS:
        assert np.array_equal(o[k], r[k]), k
    for m in range(3):
        assert np.array_equal(o["layers"][m], r["layers"][m])
    assert sum(len(x) for x in o["layers"]) == ncol * 148


def test_device_entry_feeds_the_batch_entries(cuda_device, kpp):
    import torch
    cfg, d, adv_row, xadv = inputs(24, 8)
    r = call(do.layers, cfg, d, adv_row, xadv)
    t = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)
    dv = {k: t(v) for k, v in d.items() if k != "cloud"}
    dv["cloud"] = t(d["cloud"], np.int32)
    nl = 24 * 150
    z = lambda *sh, dt=torch.float64: torch.zeros(sh, dtype=dt, device=cuda_device)
    out = dict(cb1=z(nl, 4), scal=z(nl, 13), ph_rat=z(nl, 47), air=z(nl), h2o=z(nl), cvv=z(nl, 4),
               mech=torch.full((nl,), -1, dtype=torch.int32, device=cuda_device), layers=z(3, nl, dt=torch.int64),
               count=z(3, dt=torch.int64))
    dm.layers_device(cfg, dv["u0"], dv["t"], dv["p"], dv["rho"], dv["cm3"], dv["am3"], dv["xm1"], dv["conv2"], dv["cm"],
                     dv["cloud"], dv["photol_j"], out, t(adv_row, np.int32), t(xadv), dv["s1"], dv["s3"])
    torch.cuda.synchronize()
    cnt = out["count"].cpu().numpy()
    for k in ("cb1", "scal", "ph_rat", "air", "h2o", "cvv", "mech"):
        assert np.array_equal(out[k].cpu().numpy(), r[k]), k
    for m in range(3):
        assert np.array_equal(out["layers"][m, :cnt[m]].cpu().numpy(), r["layers"][m])
    assert np.array_equal(dv["cloud"].cpu().numpy(), r["cloud"]) and np.array_equal(dv["s1"].cpu().numpy(), r["s1"])
    # the rows of a mechanism's list are what the batch entries take: gather the compact rate inputs of the aer layers
    la = out["layers"][1, :cnt[1]]
    cb1_a, scal_a = out["cb1"][la], out["scal"][la]
    assert cb1_a.shape == (cnt[1], 4) and bool((scal_a[:, 5:7].sum(dim=1) >= 1).all()) and bool((scal_a[:, 7:9] == 0).all())
    lt = out["layers"][2, :cnt[2]]
    assert bool((out["scal"][lt][:, 7:9].sum(dim=1) >= 1).all())
    # empty ensemble and refused sizes
    e = {k: v[:0] for k, v in d.items()}
    o = call(dm.layers, cfg, e, adv_row, xadv)
    assert o["mech"].shape == (0,) and all(len(x) == 0 for x in o["layers"])
    with pytest.raises(Exception):
        call(dm.layers, dict(cfg, box=True, n_bl=999), d, adv_row, xadv)
