"""ctypes binding of the CPU oracle of subkon/advec (kon_oracle.c).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko

MB = 18


class Grid(C.Structure):
    _fields_ = [("nka", C.c_int), ("nkt", C.c_int), ("a0m", C.c_double), ("dlne", C.c_double)] + [
        (n, C.POINTER(C.c_double)) for n in ("en", "rn", "b0m", "ew", "e", "dew", "rw", "qabs")] + [
        ("kw", C.POINTER(C.c_int)), ("rq", C.POINTER(C.c_double)), ("ka", C.c_int), ("reserved", C.c_int)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _grid(g):
    keep = [np.ascontiguousarray(g[n], dtype=np.float64) for n in ("en", "rn", "b0m", "ew", "e", "dew", "rw", "qabs")]
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    rq = np.ascontiguousarray(g["rq"], dtype=np.float64)
    keep += [kw, rq]
    s = Grid(g["nka"], g["nkt"], g["a0m"], g["dlne"], *[_dp(a) for a in keep[:8]],
             kw.ctypes.data_as(C.POINTER(C.c_int)), _dp(rq), g["ka"], 0)
    return s, keep


def advec(dt, u, y):
    """SUBROUTINE advec (str.f90:5321-5516) on one row.  Returns (y_new, err)."""
    L = _ko.lib()
    u = np.ascontiguousarray(u, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64).copy()
    L.kon_oracle_advec.restype = C.c_int
    err = L.kon_oracle_advec(C.c_int(len(y)), C.c_double(dt), _dp(u), _dp(y))
    return y, err


def subkon(g, dt, ffk, totr, dfdt, feualt, pp, to, tn, xm1o, xm1n, kr, nthreads=None):
    """SUBROUTINE subkon (str.f90:4987-5204) for every layer.  Returns (ffk, to, xm1o, status)."""
    L = _ko.lib()
    gs, keep = _grid(g)
    ffk = np.ascontiguousarray(ffk, dtype=np.float64).copy()
    n = ffk.shape[0]
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (totr, dfdt, feualt, pp, tn, xm1n)]
    to = np.ascontiguousarray(to, dtype=np.float64).copy()
    xm1o = np.ascontiguousarray(xm1o, dtype=np.float64).copy()
    kr = np.ascontiguousarray(kr, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    L.kon_oracle_subkon.restype = None
    L.kon_oracle_subkon(C.byref(gs), C.c_int64(n), C.c_double(dt), _dp(ffk), _dp(a[0]), _dp(a[1]), _dp(a[2]),
                        _dp(a[3]), _dp(to), _dp(a[4]), _dp(xm1o), _dp(a[5]),
                        kr.ctypes.data_as(C.POINTER(C.c_int32)), status.ctypes.data_as(C.POINTER(C.c_int32)))
    return ffk, to, xm1o, status


STATE_D = ("ff", "t", "talt", "xm1", "xm1a", "feu", "dfddt", "xm2", "dtcon", "p", "totrad")
SUMS = ("vol1_a", "vol1_d", "part_o_a", "part_o_d", "part_n_a", "part_n_d", "vol2", "pntot")


class State(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in STATE_D] + [("nar", C.POINTER(C.c_int32))] + [
        (n, C.POINTER(C.c_double)) for n in SUMS] + [("status", C.POINTER(C.c_int32))]


def layers(g, dt, chem, st):
    """The layer loop of SUBROUTINE kon (str.f90:4615-4772).  `st`: dict with ff, t, talt, xm1, xm1a,
    feu, dfddt, p, totrad, nar.  Returns a dict of the updated arrays (+ xm2, dtcon, status, sums)."""
    L = _ko.lib()
    gs, keep = _grid(g)
    n = st["ff"].shape[0]
    nka = g["nka"]
    o = {k: np.ascontiguousarray(st[k], dtype=np.float64).copy() for k in ("ff", "t", "talt", "xm1", "xm1a", "feu",
                                                                          "dfddt", "p", "totrad")}
    o["xm2"] = np.zeros(n); o["dtcon"] = np.zeros(n)
    o["nar"] = np.ascontiguousarray(st["nar"], dtype=np.int32)
    for k in SUMS[:6]:
        o[k] = np.zeros((n, nka))
    o["vol2"] = np.zeros((n, 4)); o["pntot"] = np.zeros((n, 4))
    o["status"] = np.zeros(n, dtype=np.int32)
    s = State(*[_dp(o[k]) for k in STATE_D], o["nar"].ctypes.data_as(C.POINTER(C.c_int32)),
              *[_dp(o[k]) for k in SUMS], o["status"].ctypes.data_as(C.POINTER(C.c_int32)))
    L.kon_oracle_layers.restype = None
    L.kon_oracle_layers(C.byref(gs), C.c_int64(n), C.c_double(dt), C.c_int(1 if chem else 0), C.byref(s))
    return o
