"""SUBROUTINE difc on the device (row N4, first piece; include/mistra_difc.h) vs the CPU oracle (-m gpu):
every species is one sequential recurrence in the reference's operation order, so the comparison is
bit for bit."""
import os

import numpy as np
import pytest

from mistra_b200 import difc as dm
from oracle import difc_oracle as dfo
from tests.test_difc_oracle import difp_inputs, inputs, run

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "difc_columns.npz")


def test_golden_columns(cuda_device, kpp):
    z = np.load(GOLD)
    c = {k: z[k] for k in ("atkh", "w", "am3", "detw", "deta")}
    n0 = dm.launch_count()
    outs = run(dm.difc, float(z["dt"]), c, [(z["f%d" % i], int(z["nproc"][i])) for i in range(4)])
    assert dm.launch_count() == n0 + 2                               # coefficients + one launch for all arrays
    for i, o in enumerate(outs):
        assert np.array_equal(o, z["o%d" % i])


@pytest.mark.parametrize("ncol,seed,n,nkc_l,dt", [(40, 5, 150, 4, 60.0), (7, 6, 150, 2, 10.0), (1, 7, 60, 1, 120.0),
                                                  (300, 8, 33, 4, 60.0)])
def test_synthetic_columns_vs_oracle(cuda_device, kpp, ncol, seed, n, nkc_l, dt):
    c, fields = inputs(ncol, seed, n=n, nkc_l=nkc_l)
    ref = run(dfo.difc, dt, c, fields)
    out = run(dm.difc, dt, c, fields)
    for (a, nproc), o, r in zip(fields, out, ref):
        assert np.array_equal(o, r)
        assert (o[:, 1:-1, :nproc] != a[:, 1:-1, :nproc]).any()


def test_device_entry_well_mixed_and_edges(cuda_device, kpp):
    import torch
    c, fields = inputs(25, 9, nkc_l=3)
    ref = run(dfo.difc, 60.0, c, fields)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(cuda_device)
    fd = [(t(a), p) for a, p in fields]
    dm.difc_device(60.0, t(c["atkh"]), t(c["w"]), t(c["am3"]), t(c["detw"]), t(c["deta"]), fd)
    torch.cuda.synchronize()
    for (o, _), r in zip(fd, ref):
        assert np.array_equal(o.cpu().numpy(), r)
    # a well-mixed profile stays well mixed when nothing subsides
    c["w"][:] = 0.0
    s = np.repeat(c["am3"][:, :, None], 5, axis=2) * np.arange(1.0, 6.0)
    out = run(dm.difc, 10.0, c, [(s, 5)])[0]
    assert np.allclose(out, s, rtol=1e-13)
    # empty ensemble, no fields, nothing to diffuse
    z = {k: (v[:0] if v.ndim == 2 else v) for k, v in c.items()}
    assert run(dm.difc, 10.0, z, [(s[:0], 5)])[0].shape == (0, 150, 5)
    assert run(dm.difc, 10.0, c, []) == []
    assert np.array_equal(run(dm.difc, 10.0, c, [(s, 0)])[0], s)
    with pytest.raises(ValueError):
        run(dm.difc, 10.0, c, [(s[:, :10], 5)])
    with pytest.raises(Exception):
        run(dm.difc, 10.0, c, [(s, 6)])                              # nproc > row


@pytest.mark.parametrize("ncol,seed,n,row", [(6, 3, 150, 70 * 70), (40, 4, 60, 77), (1, 5, 150, 128)])
def test_difp_vs_oracle(cuda_device, kpp, ncol, seed, n, row):
    """SUBROUTINE difp (str.f90:3137-3265): spectrum bit for bit, fsum to 1e-13 (partial sums)."""
    c, rho, ff, fsum = difp_inputs(ncol, seed, n=n, row=row)
    rf, rs = dfo.difp(60.0, c["atkh"], c["w"], rho, c["detw"], c["deta"], ff, fsum)
    of, os_ = dm.difp(60.0, c["atkh"], c["w"], rho, c["detw"], c["deta"], ff, fsum)
    assert np.array_equal(of, rf) and (of[:, 1:-1] != ff[:, 1:-1]).any()
    assert np.array_equal(os_[:, 0], fsum[:, 0]) and np.allclose(os_, rs, rtol=1e-13, atol=0)


def test_difp_device_entry_and_edges(cuda_device, kpp):
    import torch
    c, rho, ff, fsum = difp_inputs(5, 6, row=300)
    rf, rs = dfo.difp(30.0, c["atkh"], c["w"], rho, c["detw"], c["deta"], ff, fsum)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(cuda_device)
    fd, sd = t(ff), t(fsum)
    dm.difp_device(30.0, t(c["atkh"]), t(c["w"]), t(rho), t(c["detw"]), t(c["deta"]), fd, sd)
    torch.cuda.synchronize()
    assert np.array_equal(fd.cpu().numpy(), rf) and np.allclose(sd.cpu().numpy(), rs, rtol=1e-13, atol=0)
    e = dm.difp(30.0, c["atkh"][:0], c["w"][:0], rho[:0], c["detw"], c["deta"], ff[:0], fsum[:0])
    assert e[0].shape == (0, 150, 300)
    with pytest.raises(ValueError):
        dm.difp(30.0, c["atkh"], c["w"], rho[:, :10], c["detw"], c["deta"], ff, fsum)
