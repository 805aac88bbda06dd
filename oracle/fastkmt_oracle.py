"""ctypes binding of the CPU oracle of SUBROUTINE fast_k_mt_a / fast_k_mt_t and FUNCTION vterm
(fastkmt_oracle.c).  TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko


def vterm(a, t, p):
    """str.f90:2793-2864, elementwise."""
    L = _ko.lib()
    dp = C.POINTER(C.c_double)
    a, t, p = (np.ascontiguousarray(np.broadcast_to(x, np.broadcast(a, t, p).shape), dtype=np.float64) for x in (a, t, p))
    out = np.empty_like(a)
    L.vterm_oracle_vec.restype = None
    L.vterm_oracle_vec.argtypes = [C.c_int64, dp, dp, dp, dp]
    L.vterm_oracle_vec(a.size, *[x.ctypes.data_as(dp) for x in (a, t, p, out)])
    return out


def fast_k_mt(g, lex, ff, freep, t, p, cw, cm, alpha, vmean, xkmt, vt, nkc_l=4, ial=1):
    """kpp.f90:2683-2947 (= 2421-2676) for every layer.  xkmt [ncell,nkc,nspec] and vt [ncell,nkc]
    are the previous values; returns updated copies."""
    L = _ko.lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    ff, freep, t, p, cw, cm, alpha, vmean = map(f8, (ff, freep, t, p, cw, cm, alpha, vmean))
    xkmt, vt = f8(xkmt).copy(), f8(vt).copy()
    n, nkc, nspec = xkmt.shape
    lex = np.ascontiguousarray(lex, dtype=np.int32)
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    rq = f8(g["rq"])
    L.fastkmt_oracle.restype = None
    L.fastkmt_oracle.argtypes = [C.c_int64] + [C.c_int] * 8 + [ip, ip] + [dp] * 11
    L.fastkmt_oracle(n, int(g["nka"]), int(g["nkt"]), int(g["ka"]), int(ial), nkc, int(nkc_l), nspec, lex.size,
                     lex.ctypes.data_as(ip), kw.ctypes.data_as(ip),
                     *[x.ctypes.data_as(dp) for x in (rq, ff, freep, t, p, cw, cm, alpha, vmean, xkmt, vt)])
    return xkmt, vt
