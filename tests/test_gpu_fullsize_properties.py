"""Size-independent properties of the particle-grid / column kernels at the sizes bench.py runs them (-m gpu),
where the CPU oracle would take minutes: fast_k_mt is invariant under a common scaling of the spectrum and the
liquid water content; difc keeps a well-mixed profile and conserves the column integral of every mixing ratio when
nothing subsides and the boundary fluxes vanish; difp's fsum is the sum of what it wrote; the drive copies
round-trip."""
import numpy as np
import pytest

from mistra_b200 import cwrc, difc as dm, drive, fastkmt, kon
from mistra_b200.mechgen import mech as mechmod

pytestmark = pytest.mark.gpu


def test_fast_k_mt_scaling_invariance_29600_layers(cuda_device, kpp):
    import torch
    n, ns = 29600, 262
    g = kon.kon_grid()
    st = kon.synthetic_columns(g, n, seed=31, dry_fraction=0.3)
    t = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(cuda_device)
    gd = {"nka": g["nka"], "nkt": g["nkt"], "ka": g["ka"], "kw": t(g["kw"], np.int32), "e": t(g["e"]), "rq": t(g["rq"])}
    ff = t(st["ff"] * 100.0)
    feu, cloud = t(st["feu"]), torch.ones((n, 4), dtype=torch.int32, device=cuda_device)
    gen = torch.Generator(device=cuda_device).manual_seed(5)
    alpha = torch.rand((n, ns), dtype=torch.float64, device=cuda_device, generator=gen) * 0.9 + 0.01
    vmean = torch.rand((n, ns), dtype=torch.float64, device=cuda_device, generator=gen) * 600.0 + 100.0
    freep, tt, pp, lex = t(2.28e-5 * st["t"] / st["p"]), t(st["t"]), t(st["p"]), t(fastkmt.lex("aer"), np.int32)
    res = []
    for scale in (1.0, 4.0):                                     # a power of two: every term scales exactly
        f = ff * scale
        outs = [torch.empty((n, 4), dtype=torch.float64, device=cuda_device) for _ in range(4)]
        cwrc.cw_rc_device(gd, f, feu, cloud, *outs)
        cm = (outs[2] > 0).double()                              # the same bins switched on in both runs
        xk = torch.zeros((n, 4, ns), dtype=torch.float64, device=cuda_device)
        vt = torch.zeros((n, 4), dtype=torch.float64, device=cuda_device)
        fastkmt.fast_k_mt_device(gd, lex, f, freep, tt, pp, outs[1], cm, alpha, vmean, xk, vt)
        res.append((xk, vt, cm))
    torch.cuda.synchronize()
    same = (res[0][2] == res[1][2]).all(dim=1)                   # layers whose switches did not move with the scale
    assert same.float().mean().item() > 0.5
    assert torch.equal(res[0][0][same], res[1][0][same]) and torch.equal(res[0][1][same], res[1][1][same])
    assert (res[0][0] > 0).any().item() and torch.isfinite(res[0][0]).all().item()
    # species outside lex are never written
    mask = torch.ones(ns, dtype=torch.bool, device=cuda_device); mask[(lex - 1).long()] = False
    assert not res[0][0][:, :, mask].any().item()


def test_difc_difp_conservation_2000_columns(cuda_device, kpp):
    import torch
    ncol, n = 2000, 150
    c = dm.synthetic_columns(ncol, n, seed=3)
    c["w"][:] = 0.0                                               # no subsidence
    c["atkh"][:, 0] = 0.0                                         # no exchange with level 1 ...
    c["atkh"][:, n - 2] = 0.0                                     # ... nor with the fixed top level
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(cuda_device)
    d = {k: t(v) for k, v in c.items()}
    gen = torch.Generator(device=cuda_device).manual_seed(9)
    s = torch.rand((ncol, n, 484), dtype=torch.float64, device=cuda_device, generator=gen) * d["am3"][:, :, None]
    s0 = s.clone()
    well = d["am3"][:, :, None] * torch.arange(1.0, 6.0, dtype=torch.float64, device=cuda_device)
    well0 = well.clone()
    dm.difc_device(60.0, d["atkh"], d["w"], d["am3"], d["detw"], d["deta"], [(s, 363), (well, 5)])
    torch.cuda.synchronize()
    assert torch.equal(s[:, :, 363:], s0[:, :, 363:]) and torch.equal(s[:, 0], s0[:, 0]) and torch.equal(s[:, -1], s0[:, -1])
    assert torch.allclose(well, well0, rtol=1e-13, atol=0)        # a well-mixed profile stays well mixed
    # the operator is in flux form for the mixing ratio x = s / am3: sum_k x(k) detw(k) over the levels it couples
    # is conserved when the fluxes through both ends vanish
    w = (d["detw"][None, 1:n - 1] / d["am3"][:, 1:n - 1])[:, :, None]
    b0, b1 = (s0[:, 1:n - 1, :363] * w).sum(dim=1), (s[:, 1:n - 1, :363] * w).sum(dim=1)
    assert torch.allclose(b1, b0, rtol=1e-9, atol=0)
    assert (s[:, 1:n - 1, :363] != s0[:, 1:n - 1, :363]).any().item()
    # difp: fsum is the sum of the spectrum it wrote
    rho = d["am3"][:100] / 35.0
    ff = torch.rand((100, n, 4900), dtype=torch.float64, device=cuda_device, generator=gen)
    fs = torch.zeros((100, n), dtype=torch.float64, device=cuda_device)
    dm.difp_device(60.0, d["atkh"][:100].contiguous(), d["w"][:100].contiguous(), rho.contiguous(), d["detw"], d["deta"], ff, fs)
    torch.cuda.synchronize()
    assert torch.allclose(fs[:, 1:], ff[:, 1:].sum(dim=2), rtol=1e-12, atol=0) and not fs[:, 0].any().item()


def test_drive_round_trip_98000_cells(cuda_device, kpp):
    import torch
    m = mechmod.load("aer")
    gasph = [x for x in m.spc_names[:m.nvar] if x[-2:] not in ("l1", "l2")]
    mp = drive.to_device(drive.drive_map("aer", gasph[:75], gasph[75:95]), cuda_device)
    nc = 98000
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    arr = [torch.rand(sh, dtype=torch.float64, device=cuda_device, generator=gen) - 0.1
           for sh in ((nc, 75), (nc, 20), (nc, 4, 121), (nc, 4, 55))]
    ref = [a.clamp(min=0.0) for a in arr]
    lay = torch.randperm(nc, device=cuda_device, generator=gen)
    var = torch.zeros((nc, m.nvar), dtype=torch.float64, device=cuda_device)
    fix = torch.zeros((nc, m.nfix), dtype=torch.float64, device=cuda_device)
    one = torch.ones(nc, dtype=torch.float64, device=cuda_device)
    drive.gather_device(mp, lay, *arr, one, one, torch.ones((nc, 4), dtype=torch.float64, device=cuda_device), var, fix)
    drive.scatter_device(mp, lay, *arr, var, fix)
    torch.cuda.synchronize()
    for a, r in zip(arr, ref):                                    # gather + scatter = the clip of kpp_driver
        assert torch.equal(a, r)
    assert torch.equal(fix[:, 1], one) and (var != 0).any().item()
