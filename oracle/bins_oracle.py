"""ctypes binding of the CPU oracle of the 2-D bin redistribution (bins_oracle.c).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko

NKC, LSP, J2, J6 = 4, 9, 121, 55


class Grid(C.Structure):
    _fields_ = [("nka", C.c_int), ("nkt", C.c_int), ("ka", C.c_int), ("nkc_l", C.c_int),
                ("ial_first", C.c_int), ("reserved", C.c_int),
                ("kw", C.POINTER(C.c_int)), ("en", C.POINTER(C.c_double)), ("rq", C.POINTER(C.c_double))]


def _grid(g):
    """g: dict with nka, nkt, ka, nkc_l, ial_first, kw[nka] int32, en[nka], rq[nka,nkt]."""
    keep = (np.ascontiguousarray(g["kw"], dtype=np.int32), np.ascontiguousarray(g["en"], dtype=np.float64),
            np.ascontiguousarray(g["rq"], dtype=np.float64))
    s = Grid(g["nka"], g["nkt"], g["ka"], g["nkc_l"], g.get("ial_first", 1), 0,
             keep[0].ctypes.data_as(C.POINTER(C.c_int)), keep[1].ctypes.data_as(C.POINTER(C.c_double)),
             keep[2].ctypes.data_as(C.POINTER(C.c_double)))
    return s, keep


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def snapshot(g, ff, cm, sion1, sion1o=None):
    """str.f90:5916-5966.  Returns (sap[n,4], smp[n,4], sion1o[n,4,9])."""
    L = _ko.lib()
    gs, keep = _grid(g)
    ff = np.ascontiguousarray(ff, dtype=np.float64)
    n = ff.shape[0]
    cm = np.ascontiguousarray(cm, dtype=np.float64).reshape(n, NKC)
    sion1 = np.ascontiguousarray(sion1, dtype=np.float64).reshape(n, NKC, J6)
    sap = np.zeros((n, NKC)); smp = np.zeros((n, NKC))
    so = np.zeros((n, NKC, LSP)) if sion1o is None else np.ascontiguousarray(sion1o, dtype=np.float64).copy()
    L.bins_oracle_snapshot.restype = None
    L.bins_oracle_snapshot(C.byref(gs), C.c_int64(n), _dp(ff), _dp(cm), _dp(sion1), _dp(sap), _dp(smp), _dp(so))
    return sap, smp, so


def redistribute(g, ff, cm, cw, sap, smp, sion1o, sion1, sl1):
    """str.f90:5976-6134.  Returns (ff, sion1, sl1, nwarn) - inputs are not modified."""
    L = _ko.lib()
    gs, keep = _grid(g)
    ff = np.ascontiguousarray(ff, dtype=np.float64).copy()
    n = ff.shape[0]
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (cm, cw, sap, smp, sion1o)]
    sion1 = np.ascontiguousarray(sion1, dtype=np.float64).copy()
    sl1 = np.ascontiguousarray(sl1, dtype=np.float64).copy()
    nwarn = np.zeros(n, dtype=np.int32)
    L.bins_oracle_redistribute.restype = None
    L.bins_oracle_redistribute(C.byref(gs), C.c_int64(n), _dp(ff), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]),
                               _dp(a[4]), _dp(sion1), _dp(sl1), nwarn.ctypes.data_as(C.POINTER(C.c_int32)))
    return ff, sion1, sl1, nwarn
