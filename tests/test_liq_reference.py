"""henry_x, v_mean_x, st_coeff_x, equil_co_x (row N2: the per-layer tables of the liq_parm chain, include/mistra_liq.h)
against an independent evaluation of the reference's Fortran statements (tests/golden/make_liq_reference.py: Python back
end with typed REAL arithmetic; the product's generator emits C from the same text with its own back end).
CPU: the host build (libmistra_rconst.so); GPU: liq_tables_kernel.  Both configuration switches (lpJoyce14bc with
a_n2o5, lpBuxmann15alph), bins without liquid water, aer (2 bins) and tot (4 bins)."""
import os

import numpy as np
import pytest

from mistra_b200 import liq

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("henry", "vmean", "alpha", "xkef", "xkeb")


def check(got, g, flags, tol):
    for name, a in zip(NAMES, got):
        ref = g["%s_%d%d" % (name, flags[0], flags[1])]
        assert a.shape == ref.shape
        assert np.array_equal(a == 0.0, ref == 0.0), "%s: different zero pattern" % name
        nz = ref != 0.0
        rel = np.abs(a[nz] - ref[nz]) / np.abs(ref[nz])
        assert rel.max() <= tol, "%s: max rel err %.3e" % (name, rel.max())


@pytest.mark.parametrize("mech,name", [(1, "aer"), (2, "tot")])
@pytest.mark.parametrize("flags", [(0, 0), (1, 1)])
def test_host_tables_match_the_reference_statements(mech, name, flags):
    g = np.load(os.path.join(GOLD, "liq_reference_%s.npz" % name))
    got = liq.tables_host(mech, g["t"], g["conv2"], g["xgamma"], g["cw"], g["cm"], g["sion1_13_14"],
                          lpjoyce14bc=flags[0], lpbuxmann15alph=flags[1])
    check(got, g, flags, 1e-13)
    assert (got[2] <= 1.0).all() and (got[2] >= 0.0).all()         # alpha = min(1, alpha), default 0.1


@pytest.mark.gpu
@pytest.mark.parametrize("mech,name", [(1, "aer"), (2, "tot")])
@pytest.mark.parametrize("flags", [(0, 0), (1, 1)])
def test_device_tables_match_the_reference_statements(cuda_device, kpp, mech, name, flags):
    import torch
    g = np.load(os.path.join(GOLD, "liq_reference_%s.npz" % name))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
    n0 = liq.launch_count()
    got = liq.tables_device(mech, t(g["t"]), t(g["conv2"]), t(g["xgamma"]), t(g["cw"]), t(g["cm"]), t(g["sion1_13_14"]),
                            lpjoyce14bc=flags[0], lpbuxmann15alph=flags[1])
    torch.cuda.synchronize()
    assert liq.launch_count() == n0 + 1
    check([x.cpu().numpy() for x in got], g, flags, 1e-12)           # CUDA exp / sqrt vs libm: a few ulp


@pytest.mark.gpu
def test_device_tables_feed_the_rate_constants(cuda_device, kpp):
    """Ragged and empty batches; the outputs have the layout Update_RCONST_x reads (include/mistra_rconst.h)."""
    import torch
    g = np.load(os.path.join(GOLD, "liq_reference_aer.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
    full = liq.tables_device(1, t(g["t"]), t(g["conv2"]), t(g["xgamma"]))
    part = liq.tables_device(1, t(g["t"][:3]), t(g["conv2"][:3]), t(g["xgamma"][:3]))
    for a, b in zip(full, part):
        assert torch.equal(a[:3], b)
    empty = liq.tables_device(1, t(g["t"][:0]), t(g["conv2"][:0]), t(g["xgamma"][:0]))
    assert empty[0].shape == (0, 262) and empty[3].shape == (0, 2, 262)
