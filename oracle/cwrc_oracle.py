"""ctypes binding of the CPU oracle of SUBROUTINE cw_rc (cwrc_oracle.c).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko

CRYS = dict(xcryssulf=0.4, xcrysss=0.42, xdelisulf=0.7, xdeliss=0.75)


def cw_rc(g, ff, feu, cloud, ial=1, crys=None):
    """kpp.f90:2152-2414 for every layer.  Returns (rc, cw, cm, conv2), each [ncell,4]."""
    L = _ko.lib()
    cr = dict(CRYS, **(crys or {}))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    ff = np.ascontiguousarray(ff, dtype=np.float64)
    n = ff.shape[0]
    feu = np.ascontiguousarray(feu, dtype=np.float64)
    cloud = np.ascontiguousarray(cloud).astype(np.int32)
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    e = np.ascontiguousarray(g["e"], dtype=np.float64)
    rq = np.ascontiguousarray(g["rq"], dtype=np.float64)
    out = [np.zeros((n, 4)) for _ in range(4)]
    L.cwrc_oracle.restype = None
    L.cwrc_oracle.argtypes = [C.c_int64] + [C.c_int] * 4 + [C.c_double] * 4 + [ip, dp, dp, dp, dp, ip] + [dp] * 4
    L.cwrc_oracle(n, int(g["nka"]), int(g["nkt"]), int(g["ka"]), int(ial), cr["xcryssulf"], cr["xcrysss"],
                  cr["xdelisulf"], cr["xdeliss"], kw.ctypes.data_as(ip), e.ctypes.data_as(dp), rq.ctypes.data_as(dp),
                  ff.ctypes.data_as(dp), feu.ctypes.data_as(dp), cloud.ctypes.data_as(ip),
                  *[o.ctypes.data_as(dp) for o in out])
    return tuple(out)
