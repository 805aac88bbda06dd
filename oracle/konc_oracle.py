"""ctypes binding of the CPU oracle of SUBROUTINE konc (konc_oracle.c).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko

SUMS = ("vol1_a", "vol1_d", "part_o_a", "part_o_d", "part_n_a", "part_n_d")


def konc(ka, sums, vol2, pntot, sl1, sion1):
    """kpp.f90:3370-3585 for every layer.  Returns (sl1_new, sion1_new, warn[ncell,3])."""
    L = _ko.lib()
    dp = C.POINTER(C.c_double)
    s = [np.ascontiguousarray(sums[k], dtype=np.float64) for k in SUMS]
    ncell, nka = s[0].shape
    vol2 = np.ascontiguousarray(vol2, dtype=np.float64)
    pntot = np.ascontiguousarray(pntot, dtype=np.float64)
    sl1 = np.array(sl1, dtype=np.float64, order="C")
    sion1 = np.array(sion1, dtype=np.float64, order="C")
    warn = np.zeros((ncell, 3), dtype=np.int32)
    L.konc_oracle.restype = None
    L.konc_oracle.argtypes = [C.c_int64] + [C.c_int] * 4 + [dp] * 10 + [C.POINTER(C.c_int32)]
    L.konc_oracle(ncell, nka, int(ka), sl1.shape[2], sion1.shape[2], *[a.ctypes.data_as(dp) for a in s],
                  vol2.ctypes.data_as(dp), pntot.ctypes.data_as(dp), sl1.ctypes.data_as(dp),
                  sion1.ctypes.data_as(dp), warn.ctypes.data_as(C.POINTER(C.c_int32)))
    return sl1, sion1, warn
