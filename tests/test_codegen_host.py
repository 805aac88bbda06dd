"""The GENERATED device code (mistra_b200/csrc/_gen/mech_<x>.cuh) compiled for the
host with the strict arithmetic switches (-DKPP_STRICT, no FMA contraction) and
checked bit for bit against the CPU oracle's building blocks: Jac_SP + matrix
preparation, the tile-blocked LU (must equal KppDecomp's row-wise LU to the last
bit, since every entry receives the same updates in the same order), Fun and
KppSolve.  Runs without a GPU; the same sources are what nvcc compiles."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mistra_b200.mechgen import mech as mechmod
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "mistra_b200", "csrc", "_gen")
GAMMA1 = 0.43586652150845899941601945119356


@pytest.fixture(scope="module")
def gen_sources():
    subprocess.check_call(["python", "-m", "mistra_b200.mechgen.emit_cuda"], cwd=ROOT, stdout=subprocess.DEVNULL)


def _harness(tmp_path_factory, x):
    out = tmp_path_factory.mktemp("hh") / ("h%s.so" % x)
    cmd = ["g++", "-O0", "-ffp-contract=off", "-fPIC", "-shared",
           '-DMECH_HEADER="%s"' % os.path.join(GEN, "mech_%s.cuh" % x), "-DMECH_NS=mech_%s" % x,
           "-o", str(out), os.path.join(ROOT, "tests", "host", "codegen_harness.cpp")]
    subprocess.check_call(cmd)
    L = C.CDLL(str(out))
    L.h_coef_lit.restype = C.c_char_p
    L.h_jacprep.argtypes = [C.POINTER(C.c_double), C.c_double]
    return L


@pytest.mark.parametrize("mi,name", [(0, "gas"), (1, "aer"), (2, "tot")])
def test_generated_code_is_bit_identical_to_oracle_blocks(gen_sources, oracle, tmp_path_factory, mi, name):
    m = mechmod.load(name)
    L = _harness(tmp_path_factory, m.suffix)
    dp = C.POINTER(C.c_double)
    nslot, SG, SY, SR, SF, SK1 = L.h_nslot(), L.h_sg(), L.h_sy(), L.h_srct(), L.h_sfix(), L.h_sk1()
    coef = np.array([m.literal_value(L.h_coef_lit(i).decode(), 1) for i in range(L.h_ncoef())])
    L.h_set_coef(coef.ctypes.data_as(dp))
    var, fix, rc = util.random_cells(name, 3, 4321 + mi)
    for c in range(3):
        w = np.full(nslot * 32, np.nan)               # poisoned: reading a slot nobody wrote shows up
        W = w.reshape(nslot, 32)
        lane = 5 * c                                  # any lane of the interleaved workspace
        W[SY:SY + m.nvar, lane] = var[c]
        W[SF:SF + m.nfix, lane] = fix[c]
        W[SR:SR + m.nreact, lane] = rc[c]
        wp = C.cast(w.ctypes.data + 8 * lane, dp)
        ghinv = 1.0 / (10.0 ** (-2 - c) * GAMMA1)
        assert L.h_jacprep(wp, ghinv) == 0
        G = -oracle.jac(mi, var[c], fix[c], rc[c], f32=1)
        G[m.diag[:m.nvar]] += ghinv
        got = W[SG:SG + m.lu_nonzero, lane]
        # structural fill-in (Jac_SP's zeros) is neither stored by jacprep nor loaded by the LU:
        # those slots stay poisoned until the factorisation writes them
        fill = np.array([not m.jvs[nz] for nz in range(m.lu_nonzero)])
        fill[m.diag[:m.nvar]] = False
        assert np.isnan(got[fill]).all() and (G[fill] == 0).all()
        assert np.array_equal(got[~fill], G[~fill])
        L.h_decomp(wp)
        LU, ier = oracle.decomp(mi, G)
        assert ier == 0 and np.array_equal(W[SG:SG + m.lu_nonzero, lane], LU)
        L.h_fun0(wp)
        f = oracle.fun(mi, var[c], fix[c], rc[c], f32=1)
        assert np.array_equal(W[SK1:SK1 + m.nvar, lane], f)
        L.h_solve1(wp)
        assert np.array_equal(W[SK1:SK1 + m.nvar, lane], oracle.solve(mi, LU, f))
        other = np.delete(W, lane, axis=1)
        assert np.isnan(other).all()                  # nothing written outside the lane
