"""Gather / scatter halves of the *_drive routines on the device (rows a14 / a15; include/mistra_drive.h) vs their
numpy restatement (-m gpu, bit for bit), and a chemistry step that never leaves the device - gather ->
INTEGRATE_a -> scatter on the model arrays - against the same step through the host entries."""
import numpy as np
import pytest

from mistra_b200 import drive, synthetic
from mistra_b200.mechgen import mech as mechmod
from oracle import drive_oracle as dro
from tests.test_drive_maps import J1, J5, J2, J6, NKC, name_lists, state

pytestmark = pytest.mark.gpu


def dev(cuda_device):
    import torch
    return lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)


@pytest.mark.parametrize("mech,nl,nc,f32", [("aer", 300, 170, True), ("tot", 64, 64, False), ("gas", 50, 1, True)])
def test_gather_scatter_vs_oracle(cuda_device, kpp, mech, nl, nc, f32):
    import torch
    t = dev(cuda_device)
    gn, rn = name_lists(mech, seed=nl)
    mp = drive.drive_map(mech, gn, rn, allow_unmatched=True)
    md = drive.to_device(mp, cuda_device)
    m = mechmod.load(mech)
    st = state(nl, 3)
    r = np.random.default_rng(5)
    layer = r.permutation(nl)[:nc].astype(np.int64)
    air, h2o = r.uniform(30, 45, nc), r.uniform(0.1, 0.6, nc)
    cvv = np.where(r.uniform(size=(nc, 4)) < 0.3, 0.0, r.uniform(1e3, 1e6, (nc, 4)))
    var0, fix0 = r.uniform(1, 2, (nc, m.nvar)), r.uniform(1, 2, (nc, m.nfix))
    ref = dro.gather(mp, layer, st["s1"], st["s3"], st["sl1"], st["sion1"], air, h2o, cvv, var0, fix0, f32_literals=f32)
    d = {k: t(v) for k, v in st.items()}
    var, fix = t(var0), t(fix0)
    n0 = drive.launch_count()
    drive.gather_device(md, t(layer, np.int64), d["s1"], d["s3"], d["sl1"], d["sion1"], t(air), t(h2o), t(cvv), var, fix,
                        f32_literals=f32)
    torch.cuda.synchronize()
    assert drive.launch_count() == n0 + 1
    for o, x in zip((var, fix, d["sl1"], d["sion1"]), ref):
        assert np.array_equal(o.cpu().numpy(), x)
    assert np.array_equal(d["s1"].cpu().numpy(), st["s1"])
    # scatter other values back
    v2, f2 = r.uniform(-0.5, 2, (nc, m.nvar)), r.uniform(-0.5, 2, (nc, m.nfix))
    ref2 = dro.scatter(mp, layer, st["s1"], st["s3"], ref[2], ref[3], v2, f2)
    drive.scatter_device(md, t(layer, np.int64), d["s1"], d["s3"], d["sl1"], d["sion1"], t(v2), t(f2))
    torch.cuda.synchronize()
    for k, x in zip(("s1", "s3", "sl1", "sion1"), ref2):
        assert np.array_equal(d[k].cpu().numpy(), x)
    # empty batch and bad shapes
    e = t(np.zeros(0), np.int64)
    drive.gather_device(md, e, d["s1"], d["s3"], d["sl1"], d["sion1"], t(air[:0]), t(h2o[:0]), t(cvv[:0]), var[:0], fix[:0])
    with pytest.raises(ValueError):
        drive.gather_device(md, t(layer, np.int64), d["s1"], d["s3"], d["sl1"], d["sion1"], t(air), t(h2o), t(cvv),
                            var[:, :5].contiguous(), fix)


def test_chemistry_step_stays_on_the_device(cuda_device, kpp):
    import torch
    t = dev(cuda_device)
    ens = synthetic.AerEnsemble(2, seed=7)                     # 196 aer cells with consistent VAR, FIX, RCONST
    m = mechmod.load("aer")
    nc = ens.ncell
    gn, rn = name_lists("aer", seed=1)
    mp = drive.drive_map("aer", gn, rn, allow_unmatched=True)
    md = drive.to_device(mp, cuda_device)
    nl = nc + 30
    layer = np.random.default_rng(2).permutation(nl)[:nc].astype(np.int64)
    junk = state(nl, 8)
    # model arrays that hold the ensemble's concentrations (the oracle scatter builds them)
    s1, s3, sl1, sion1 = dro.scatter(mp, layer, junk["s1"], junk["s3"], junk["sl1"], junk["sion1"], ens.var, ens.fix,
                                     clip_negative=True)
    air = ens.fix[:, mp["indf_o2"] - 1] / float(np.float32(0.21))
    h2o = ens.fix[:, mp["indf_h2o"] - 1]
    cvv = np.zeros((nc, 4))
    for b in (0, 1):
        f = ens.fix[:, mp["indf_h2ol"][b] - 1]
        cvv[:, b] = np.where(f > 0, float(np.float32(55.55)) / np.where(f > 0, f, 1.0), 0.0)
    rconst = ens.rconst()
    # ---- host path: numpy gather -> mistra_kpp_integrate (host buffers) -> numpy scatter ----
    var_h, fix_h, sl1_h, sion1_h = dro.gather(mp, layer, s1, s3, sl1, sion1, air, h2o, cvv, ens.var, ens.fix)
    out = kpp.integrate(1, rconst, fix_h, var_h)
    ref = dro.scatter(mp, layer, s1, s3, sl1_h, sion1_h, out[0], fix_h)
    # ---- device path ----
    d = dict(s1=t(s1), s3=t(s3), sl1=t(sl1), sion1=t(sion1))
    var, fix, ld = t(ens.var), t(ens.fix), t(layer, np.int64)
    drive.gather_device(md, ld, d["s1"], d["s3"], d["sl1"], d["sion1"], t(air), t(h2o), t(cvv), var, fix)
    kpp.integrate_device(1, t(rconst), fix, var)
    drive.scatter_device(md, ld, d["s1"], d["s3"], d["sl1"], d["sion1"], var, fix)
    torch.cuda.synchronize()
    for k, x in zip(("s1", "s3", "sl1", "sion1"), ref):
        assert np.array_equal(d[k].cpu().numpy(), x), k
    assert (ref[0][layer] != s1[layer]).any()                  # the chemistry moved something
