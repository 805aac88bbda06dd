"""ctypes binding of the CPU oracle of SUBROUTINE difc (difc_oracle.c).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko


def difc(dt, atkh, w, am3, detw, deta, fields):
    """str.f90:3271-3445 for every column.  fields: list of (array [ncol,n,row], nproc); returns the
    updated copies."""
    L = _ko.lib()
    dp = C.POINTER(C.c_double)
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    atkh, w, am3, detw, deta = map(f8, (atkh, w, am3, detw, deta))
    ncol, n = atkh.shape
    outs = [f8(a).copy() for a, _ in fields]
    ptrs = (dp * max(1, len(outs)))(*[o.ctypes.data_as(dp) for o in outs])
    row = np.array([o.shape[2] for o in outs], dtype=np.int32)
    npr = np.array([p for _, p in fields], dtype=np.int32)
    L.difc_oracle.restype = None
    L.difc_oracle.argtypes = [C.c_int64, C.c_int, C.c_double, dp, dp, dp, dp, dp, C.c_int, C.POINTER(dp),
                              C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.difc_oracle(ncol, n, float(dt), *[x.ctypes.data_as(dp) for x in (atkh, w, am3, detw, deta)], len(outs), ptrs,
                  row.ctypes.data_as(C.POINTER(C.c_int32)), npr.ctypes.data_as(C.POINTER(C.c_int32)))
    return outs


def difp(dt, atkh, w, rho, detw, deta, ff, fsum):
    """str.f90:3137-3265 for every column.  ff [ncol,n,row]; returns updated copies (ff, fsum)."""
    L = _ko.lib()
    dp = C.POINTER(C.c_double)
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    atkh, w, rho, detw, deta = map(f8, (atkh, w, rho, detw, deta))
    ff, fsum = f8(ff).copy(), f8(fsum).copy()
    ncol, n, row = ff.shape
    L.difp_oracle.restype = None
    L.difp_oracle.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_double] + [dp] * 7
    L.difp_oracle(ncol, n, row, float(dt), *[x.ctypes.data_as(dp) for x in (atkh, w, rho, detw, deta, ff, fsum)])
    return ff, fsum
