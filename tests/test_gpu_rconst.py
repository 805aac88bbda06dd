"""Update_RCONST_x on the device (row N1) vs the host library on the same inputs (-m gpu).
Same generated expressions and rate-law code; exp/pow/log10 from the CUDA math library:
1e-12 relative.  The device rate constants then drive the integrator without leaving HBM."""
import numpy as np
import pytest

from mistra_b200 import rconst as rc
from mistra_b200 import synthetic

pytestmark = pytest.mark.gpu
FIELDS = ("yhenry", "yxkmt", "ykef", "ykeb", "yxkmtd", "yxeq", "ycw", "ycwd")


def device_rconst(ens, dev, f32):
    import torch
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    kw = {k: t(getattr(ens, k, None)) for k in FIELDS}
    out = rc.update_rconst_device(ens.mech, t(ens.cb1), t(ens.scal), t(ens.ph_rat), t(ens.conc()),
                                  f32_literals=f32, **kw)
    return out


@pytest.mark.parametrize("cls,ncol", [(synthetic.GasEnsemble, 3), (synthetic.AerEnsemble, 2), (synthetic.TotEnsemble, 1)])
@pytest.mark.parametrize("f32", [1, 0])
def test_device_rconst_matches_host(cuda_device, kpp, cls, ncol, f32):
    ens = cls(ncol, f32_literals=f32)
    host = ens.rconst()
    dev = device_rconst(ens, cuda_device, f32).cpu().numpy()
    assert dev.shape == host.shape
    nz = host != 0
    assert np.array_equal(dev == 0, ~nz)                       # the same switches zero the same reactions
    rel = np.abs(dev[nz] - host[nz]) / np.abs(host[nz])
    assert rel.max() <= 1e-12, rel.max()


def test_device_rconst_feeds_the_integrator(cuda_device, kpp):
    import torch
    ens = synthetic.AerEnsemble(2)
    d_rc = device_rconst(ens, cuda_device, 1)
    var = torch.from_numpy(ens.var).to(cuda_device)
    fix = torch.from_numpy(np.ascontiguousarray(ens.fix)).to(cuda_device)
    ierr = torch.zeros(ens.ncell, dtype=torch.int32, device=cuda_device)
    stats = torch.zeros((ens.ncell, 8), dtype=torch.int32, device=cuda_device)
    n0 = kpp.launch_count()
    kpp.integrate_device(1, d_rc, fix, var, ierr=ierr, stats=stats)
    torch.cuda.synchronize()
    ref, ierr_h, stats_h, _, _ = kpp.integrate(1, ens.rconst(), ens.fix, ens.var)
    assert kpp.launch_count() >= n0 + 2 and (ierr.cpu().numpy() == 1).all()
    out = var.cpu().numpy()
    rel = np.abs(out - ref) / (np.maximum(np.abs(out), np.abs(ref)) + 1.66e-21)
    assert rel.max() <= 1e-3                                  # north-star tolerance; same steps in nearly all cells
    assert (stats.cpu().numpy()[:, 2] == stats_h[:, 2]).mean() > 0.9


def test_bad_arguments(cuda_device, kpp):
    import torch
    ens = synthetic.GasEnsemble(1)
    with pytest.raises(ValueError):
        rc.update_rconst_device(0, torch.zeros((ens.ncell, 3), dtype=torch.float64, device=cuda_device),
                                torch.zeros((ens.ncell, 13), dtype=torch.float64, device=cuda_device),
                                torch.zeros((ens.ncell, 47), dtype=torch.float64, device=cuda_device),
                                torch.zeros((ens.ncell, 105), dtype=torch.float64, device=cuda_device))
