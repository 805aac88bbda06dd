"""The particle-grid kernels chained on one stream without a host copy (-m gpu): condensation (subkon) ->
cw_rc -> fast_k_mt_a -> bin snapshot on the spectrum that stays in device memory, as INTEGRATION.md
describes for kon -> liq_parm -> stem_kpp.  The chained result must equal, bit for bit, the same entries
called one by one through their host-buffer forms, and agree with the CPU oracles fed the same spectrum."""
import numpy as np
import pytest

from mistra_b200 import bins, cwrc, fastkmt, kon
from oracle import cwrc_oracle as cwo
from oracle import fastkmt_oracle as fko

pytestmark = pytest.mark.gpu


def test_device_resident_chain(cuda_device, kpp):
    import torch
    n, ns = 600, 262
    g = kon.kon_grid()
    d = kon.synthetic_layers(g, n, seed=21)
    r = np.random.default_rng(22)
    d["ffk"] = d["ffk"] * 10.0 ** r.uniform(0, 3, (n, 1, 1))            # up to cloudy water contents
    keys = ("ffk", "totr", "dfdt", "feualt", "pp", "to", "tn", "xm1o", "xm1n", "kr")
    cloud = (r.uniform(size=(n, 4)) < 0.5).astype(np.int32)
    alpha = 10.0 ** r.uniform(-4, 0, (n, ns))
    vmean = r.uniform(100.0, 700.0, (n, ns))
    freep = 2.28e-5 * d["tn"] / d["pp"]
    lex = fastkmt.lex("aer")
    xk0, vt0 = np.zeros((n, 4, ns)), np.zeros((n, 4))
    sion1 = r.uniform(0.0, 1.0e-3, (n, 4, 55))

    # ---- one by one through the host-buffer entries ----
    ff1, to1, xm1, st1 = kon.subkon(g, 10.0, *[d[k] for k in keys])
    rc1, cw1, cm1, cv1 = cwrc.cw_rc(g, ff1, d["feualt"], cloud)
    xk1, vt1 = fastkmt.fast_k_mt(g, lex, ff1, freep, d["tn"], d["pp"], cw1, cm1, alpha, vmean, xk0, vt0)
    bg = bins.particle_grid()
    sap1, smp1, so1 = bins.snapshot(bg, ff1, cm1, sion1)

    # ---- chained on the device ----
    t = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)
    dd = {k: t(d[k], np.int32 if k == "kr" else np.float64) for k in keys}
    stat = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    gd = {"nka": g["nka"], "nkt": g["nkt"], "ka": g["ka"], "kw": t(g["kw"], np.int32), "e": t(g["e"]), "rq": t(g["rq"])}
    outs = [torch.empty((n, 4), dtype=torch.float64, device=cuda_device) for _ in range(4)]
    xk, vt = t(xk0), t(vt0)
    # the snapshot assigns sion1o only for bins with chemistry (str.f90:5950-5960); start from zeros like the host form
    sap, smp = (torch.zeros((n, 4), dtype=torch.float64, device=cuda_device) for _ in range(2))
    so = torch.zeros((n, 4, 9), dtype=torch.float64, device=cuda_device)
    feu_d, cloud_d, lex_d, fr_d, al_d, vm_d, si_d = (t(d["feualt"]), t(cloud, np.int32), t(lex, np.int32), t(freep),
                                                      t(alpha), t(vmean), t(sion1))
    torch.cuda.synchronize()
    kon.subkon_device(g, 10.0, *[dd[k] for k in keys], status=stat)
    cwrc.cw_rc_device(gd, dd["ffk"], feu_d, cloud_d, *outs)
    fastkmt.fast_k_mt_device(gd, lex_d, dd["ffk"], fr_d, dd["tn"], dd["pp"], outs[1], outs[2], al_d, vm_d, xk, vt)
    bins.snapshot_device(bg, dd["ffk"], outs[2], si_d, sap, smp, so)
    torch.cuda.synchronize()

    h = lambda x: x.cpu().numpy()
    assert np.array_equal(h(dd["ffk"]), ff1) and np.array_equal(h(stat), st1)
    for a, b in zip(outs, (rc1, cw1, cm1, cv1)):
        assert np.array_equal(h(a), b)
    assert np.array_equal(h(xk), xk1) and np.array_equal(h(vt), vt1)
    assert np.array_equal(h(sap), sap1) and np.array_equal(h(smp), smp1) and np.array_equal(h(so), so1)
    assert (cm1 > 0).any() and (xk1 != 0).any()

    # ---- and the oracles on the spectrum the device produced ----
    ro = cwo.cw_rc(g, ff1, d["feualt"], cloud)
    for a, b in zip((rc1, cw1, cm1, cv1), ro):
        assert np.array_equal(a == 0, b == 0) and np.allclose(a, b, rtol=1e-13, atol=0)
    xo, vo = fko.fast_k_mt(g, lex, ff1, freep, d["tn"], d["pp"], cw1, cm1, alpha, vmean, xk0, vt0)
    assert np.allclose(xk1, xo, rtol=1e-13, atol=0) and np.allclose(vt1, vo, rtol=1e-13, atol=0)
