"""Host-side rate constants: ctypes mirror of include/mistra_rconst.h
(batched Update_RCONST_g/_a/_t, /root/reference/src/gas.f:275, aer.f:304,
tot.f:1040).  Used by the synthetic-ensemble generator and the tests; a Fortran
caller keeps its own Update_RCONST_x."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .mechgen import mech as mechmod

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NPHRXN = 47
NKC = {0: 2, 1: 2, 2: 4}


class RateInputs(C.Structure):
    _fields_ = [("ncell", C.c_int64)] + [(n, C.POINTER(C.c_double)) for n in (
        "cb1", "scal", "ph_rat", "conc", "yhenry", "yxkmt", "ykef", "ykeb", "yxkmtd", "yxeq",
        "ycw", "ycwd")] + [("f32_literals", C.c_int32), ("reserved", C.c_int32)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libmistra_rconst.so")
        if not os.path.exists(so):
            raise RuntimeError("libmistra_rconst.so missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(so)
        L.mistra_rconst_update.argtypes = [C.c_int, C.POINTER(RateInputs), C.POINTER(C.c_double), C.c_int]
        L.mistra_rconst_spc_index.argtypes = [C.c_int, C.c_char_p]
        _LIB = L
    return _LIB


def spc_index(mech, name):
    return lib().mistra_rconst_spc_index(mech, name.encode())


def update_rconst(mech, cb1, scal, ph_rat, conc, yhenry=None, yxkmt=None, ykef=None, ykeb=None,
                  yxkmtd=None, yxeq=None, ycw=None, ycwd=None, f32_literals=1, nthreads=0):
    """All arrays float64, one row per cell (see mistra_rconst.h). Returns [ncell][NREACT]."""
    m = mechmod.load(mechmod.MECH_NAMES[mech])
    nspec = m.nvar + m.nfix
    keep = []

    def arr(a, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == shape, (a.shape, shape)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double))

    conc = np.ascontiguousarray(conc, dtype=np.float64)
    n = conc.shape[0]
    nkc = NKC[mech]
    ri = RateInputs()
    ri.ncell = n
    ri.cb1 = arr(cb1, (n, 4))
    ri.scal = arr(scal, (n, 13))
    ri.ph_rat = arr(ph_rat, (n, NPHRXN))
    ri.conc = arr(conc, (n, nspec))
    ri.yhenry = arr(yhenry, (n, nspec))
    ri.yxkmt = arr(yxkmt, (n, nkc, nspec))
    ri.ykef = arr(ykef, (n, nkc, nspec))
    ri.ykeb = arr(ykeb, (n, nkc, nspec))
    ri.yxkmtd = arr(yxkmtd, (n, 2, nspec))
    ri.yxeq = arr(yxeq, (n, nspec))
    ri.ycw = arr(ycw, (n, nkc))
    ri.ycwd = arr(ycwd, (n, 2))
    ri.f32_literals = int(f32_literals)
    out = np.empty((n, m.nreact), dtype=np.float64)
    if nthreads <= 0:
        nthreads = min(os.cpu_count() or 1, 32)
    rc = lib().mistra_rconst_update(mech, C.byref(ri), out.ctypes.data_as(C.POINTER(C.c_double)), nthreads)
    if rc != 0:
        raise ValueError("mistra_rconst_update failed (%d)" % rc)
    return out


def update_rconst_device(mech, cb1, scal, ph_rat, conc, out=None, yhenry=None, yxkmt=None, ykef=None, ykeb=None,
                         yxkmtd=None, yxeq=None, ycw=None, ycwd=None, f32_literals=1, stream=None):
    """Update_RCONST_x on the device (include/mistra_rconst_cuda.h): all arrays are contiguous
    float64 CUDA tensors on the current device, one row per cell; returns the [ncell][NREACT]
    tensor (`out` if given).  Asynchronous on torch's current stream."""
    import torch
    from . import kpp
    L = kpp.library()
    L.mistra_rconst_update_device.argtypes = [C.c_int, C.POINTER(RateInputs), C.c_void_p, C.c_void_p]
    m = mechmod.load(mechmod.MECH_NAMES[mech])
    nspec, nkc, n = m.nvar + m.nfix, NKC[mech], conc.shape[0]

    def ptr(t, shape):
        if t is None:
            return None
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64 and tuple(t.shape) == shape):
            raise ValueError("update_rconst_device: need contiguous CUDA float64 %s, got %s" % (shape, tuple(t.shape)))
        return C.cast(t.data_ptr(), C.POINTER(C.c_double))
    ri = RateInputs()
    ri.ncell = n
    ri.cb1 = ptr(cb1, (n, 4)); ri.scal = ptr(scal, (n, 13)); ri.ph_rat = ptr(ph_rat, (n, NPHRXN))
    ri.conc = ptr(conc, (n, nspec)); ri.yhenry = ptr(yhenry, (n, nspec)); ri.yxkmt = ptr(yxkmt, (n, nkc, nspec))
    ri.ykef = ptr(ykef, (n, nkc, nspec)); ri.ykeb = ptr(ykeb, (n, nkc, nspec)); ri.yxkmtd = ptr(yxkmtd, (n, 2, nspec))
    ri.yxeq = ptr(yxeq, (n, nspec)); ri.ycw = ptr(ycw, (n, nkc)); ri.ycwd = ptr(ycwd, (n, 2))
    ri.f32_literals = int(f32_literals)
    if out is None:
        out = torch.empty((n, m.nreact), dtype=torch.float64, device=conc.device)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_rconst_update_device(mech, C.byref(ri), C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
    return out
