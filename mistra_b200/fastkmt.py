"""Host-side mirror of SUBROUTINE fast_k_mt_a / fast_k_mt_t (/root/reference/src/kpp.f90:2683-2947,
2421-2676) over the C ABI of include/mistra_fastkmt.h: mass-transfer coefficients xkmt of the nx = 50
exchanged species and sedimentation velocity vt of the chemistry bins, integrated over the particle
spectrum ff.  CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kpp
from .mechgen import mech as mechmod

# DATA lex / ... / of fast_k_mt_a and fast_k_mt_t (kpp.f90:2794-2803, 2532-2541): the exchanged species
LEX_NAMES = ("NO2", "HNO3", "NH3", "SO2", "H2SO4", "O3", "ACO2", "HCHO", "H2O2", "HONO", "HCl", "N2O5", "HNO4",
             "NO3", "OH", "HO2", "MO2", "CO2", "O2", "ROOH", "HOCl", "Cl2", "HBr", "HOBr", "Br2", "BrCl", "DMSO",
             "ClNO3", "BrNO3", "CH3SO3H", "DMS", "CH3SO2H", "DMSO2", "HOI", "IO", "I2", "ICl", "IBr", "OIO", "INO2",
             "INO3", "HI", "I2O2", "HIO3", "NO", "ACTA", "CH3OH", "C2H5OH", "XOR", "SOR")


def lex(mech_name):
    """1-based KPP indices (ind_X of aer_Parameters.h / tot_Parameters.h) of the exchanged species."""
    m = mechmod.load(mech_name)
    pos = {s: i + 1 for i, s in enumerate(m.spc_names)}
    return np.array([pos[s] for s in LEX_NAMES], dtype=np.int32)


class FastkmtArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nka", "nkt", "ka", "ial", "nkc", "nkc_l", "nspec", "nx")] + [
        (n, C.c_void_p) for n in ("lex", "kw", "rq", "ff", "freep", "t", "p", "cw", "cm", "alpha", "vmean",
                                  "xkmt", "vt")]


def _lib():
    L = kpp.library()
    L.mistra_fastkmt.argtypes = [C.c_int64, C.POINTER(FastkmtArgs), C.c_void_p]
    L.mistra_fastkmt_device.argtypes = [C.c_int64, C.POINTER(FastkmtArgs), C.c_void_p]
    L.mistra_fastkmt_launch_count.restype = C.c_int64
    return L


def fast_k_mt(g, lex, ff, freep, t, p, cw, cm, alpha, vmean, xkmt, vt, nkc_l=4, ial=1):
    """HOST numpy arrays: g = grid dict with nka, nkt, ka, kw, rq (kon.kon_grid()); lex [nx] 1-based
    species indices; ff [ncell,nka,nkt]; freep, t, p [ncell]; cw, cm [ncell,nkc]; alpha, vmean
    [ncell,nspec]; xkmt [ncell,nkc,nspec] and vt [ncell,nkc] hold the previous values (the reference
    leaves entries of bins without chemistry untouched).  Returns updated copies (xkmt, vt)."""
    L = _lib()
    nka, nkt = int(g["nka"]), int(g["nkt"])
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    ff, freep, t, p, cw, cm, alpha, vmean = map(f8, (ff, freep, t, p, cw, cm, alpha, vmean))
    xkmt, vt = f8(xkmt).copy(), f8(vt).copy()
    if xkmt.ndim != 3:
        raise ValueError("fast_k_mt: xkmt must be [ncell,nkc,nspec]")
    n, nkc, nspec = xkmt.shape
    if ff.shape != (n, nka, nkt):
        raise ValueError("fast_k_mt: ff must be [ncell,nka,nkt]")
    for name, x, shp in (("freep", freep, (n,)), ("t", t, (n,)), ("p", p, (n,)), ("cw", cw, (n, nkc)),
                         ("cm", cm, (n, nkc)), ("alpha", alpha, (n, nspec)), ("vmean", vmean, (n, nspec)),
                         ("vt", vt, (n, nkc))):
        if x.shape != shp:
            raise ValueError("fast_k_mt: %s must be %s" % (name, shp))
    lex = np.ascontiguousarray(lex, dtype=np.int32)
    if lex.ndim != 1 or lex.size < 1 or lex.min() < 1 or lex.max() > nspec:
        raise ValueError("fast_k_mt: lex must hold 1-based species indices <= nspec")
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    rq = f8(g["rq"])
    a = FastkmtArgs(nka, nkt, int(g["ka"]), int(ial), nkc, int(nkc_l), nspec, lex.size,
                    *[x.ctypes.data for x in (lex, kw, rq, ff, freep, t, p, cw, cm, alpha, vmean, xkmt, vt)])
    kpp._check(L, L.mistra_fastkmt(n, C.byref(a), None))
    return xkmt, vt


def fast_k_mt_device(g_dev, lex, ff, freep, t, p, cw, cm, alpha, vmean, xkmt, vt, nkc_l=4, ial=1, stream=None):
    """Same on torch CUDA tensors of the current device; xkmt and vt are updated in place.  g_dev:
    dict with nka, nkt, ka and CUDA tensors kw (int32 [nka]), rq ([nka,nkt]); lex int32 CUDA tensor.
    Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib()
    nka, nkt = int(g_dev["nka"]), int(g_dev["nkt"])
    n, nkc, nspec = xkmt.shape

    def ok(x, shape, dt=torch.float64):
        if not (x.is_cuda and x.is_contiguous() and x.dtype == dt and tuple(x.shape) == shape):
            raise ValueError("fast_k_mt_device: need contiguous CUDA %s %s" % (dt, shape))
        return x.data_ptr()
    a = FastkmtArgs(nka, nkt, int(g_dev["ka"]), int(ial), nkc, int(nkc_l), nspec, int(lex.numel()),
                    ok(lex, (lex.numel(),), torch.int32), ok(g_dev["kw"], (nka,), torch.int32),
                    ok(g_dev["rq"], (nka, nkt)), ok(ff, (n, nka, nkt)), ok(freep, (n,)), ok(t, (n,)), ok(p, (n,)),
                    ok(cw, (n, nkc)), ok(cm, (n, nkc)), ok(alpha, (n, nspec)), ok(vmean, (n, nspec)),
                    ok(xkmt, (n, nkc, nspec)), ok(vt, (n, nkc)))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_fastkmt_device(n, C.byref(a), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_fastkmt_launch_count())
