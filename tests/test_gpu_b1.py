"""Boundary B1 on the GPU: the per-cell shims integrate_g_ / integrate_a_ / integrate_t_ of libmistra_kpp_f77.so, called
the way a Fortran host calls INTEGRATE_x (gas.f:173, aer.f:217, tot.f:604): state in the COMMON blocks /GDATA_x/, which
tests/host/libb1_host.so DEFINES (the shim finds them with dlsym, or through mistra_kpp_f77_bind).  Checked against the
golden vectors minted from the CPU oracle, incl. TIN = Texit, STEPMIN = Hexit and the tolerances INTEGRATE_x sets."""
import ctypes as C
import os

import numpy as np
import pytest

from mistra_b200.mechgen import mech as mechmod
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host(kpp, cuda_device):
    H = C.CDLL(os.path.join(ROOT, "tests", "host", "libb1_host.so"), mode=C.RTLD_GLOBAL)
    H.b1_integrate.argtypes = [C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp]
    H.b1_latency_us.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_int]
    H.b1_latency_us.restype = C.c_double
    return H


def ptr(a):
    return a.ctypes.data_as(dp)


@pytest.mark.parametrize("name", ["gas_cells", "aer_cells", "tot_cells"])
def test_shim_matches_golden_cells(host, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    mech = int(g["mech"])
    ok = np.nonzero(g["ierr"] == 1)[0][:4]
    for c in ok:
        var = np.ascontiguousarray(g["var"][c])
        fix = np.ascontiguousarray(g["fix"][c])
        rc = np.ascontiguousarray(g["rconst"][c])
        out = np.empty_like(var)
        texit, stepmin, tol = np.zeros(1), np.zeros(1), np.zeros(2)
        assert host.b1_integrate(mech, ptr(var), ptr(fix), ptr(rc), ptr(out), 0.0, 10.0, ptr(texit), ptr(stepmin), ptr(tol)) == 0
        assert util.rel_err(out, g["var_out"][c]).max() <= util.RTOL
        assert texit[0] == 10.0                                           # TIN = Texit, gas.f:769
        assert np.isclose(stepmin[0], g["hexit"][c], rtol=1e-4)           # STEPMIN = Hexit, gas.f:770
        assert tol[0] == 1.0e-25 and tol[1] == 1.0e-3                     # gas.f:745-746


def test_bound_image_wins_over_the_symbol(host, kpp):
    """mistra_kpp_f77_bind: the shim works on the image it was handed (a host whose COMMON blocks are not in the
    dynamic symbol table registers them once at start-up)."""
    F = C.CDLL(os.path.join(ROOT, "mistra_b200", "libmistra_kpp_f77.so"))
    g = np.load(os.path.join(GOLD, "gas_cells.npz"))
    m = mechmod.load("gas")
    c = int(np.nonzero(g["ierr"] == 1)[0][0])
    img = np.zeros(644)                                                   # C(105) RCONST(331) TIME DT ATOL(102) RTOL(102) STEPMIN STEPMAX
    img[:m.nvar] = g["var"][c]
    img[m.nvar:m.nvar + m.nfix] = g["fix"][c]
    img[105:105 + m.nreact] = g["rconst"][c]
    F.mistra_kpp_f77_bind.argtypes = [C.c_int, C.c_void_p]
    F.mistra_kpp_f77_bind(0, img.ctypes.data)
    try:
        tin, tout = C.c_double(0.0), C.c_double(10.0)
        F.integrate_g_(C.byref(tin), C.byref(tout))
        assert tin.value == 10.0
        assert util.rel_err(img[:m.nvar], g["var_out"][c]).max() <= util.RTOL
        assert img[105 + 331 + 2] == 1.0e-25 and img[105 + 331 + 2 + 102] == 1.0e-3
    finally:
        F.mistra_kpp_f77_bind(0, None)


def test_latency_loop_runs(host):
    g = np.load(os.path.join(GOLD, "aer_cells.npz"))
    c = int(np.nonzero(g["ierr"] == 1)[0][0])
    var, fix, rc = (np.ascontiguousarray(g[k][c]) for k in ("var", "fix", "rconst"))
    us = host.b1_latency_us(1, ptr(var), ptr(fix), ptr(rc), 20, 1)
    assert 0.0 < us < 1e6
