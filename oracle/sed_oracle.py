"""ctypes binding of the CPU oracle of sedp / sedl / sedc / advsed0 / advsed1 / vterm (sed_oracle.c).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import ctypes as C

import numpy as np

from . import kpp_oracle as _ko

_dp = C.POINTER(C.c_double)
_f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
_p = lambda x: x.ctypes.data_as(_dp)


def vterm(a, t, p):
    """str.f90:2793-2864, element-wise."""
    L = _ko.lib()
    L.sed_vterm.restype = C.c_double
    L.sed_vterm.argtypes = [C.c_double] * 3
    return np.array([L.sed_vterm(float(x), float(y), float(z)) for x, y, z in zip(a, t, p)])


def advsed(scheme, c, y):
    """advsed0 (scheme 0, str.f90:5522-5579) or advsed1 (scheme 1, 5585-5691) on one profile; returns the new y."""
    L = _ko.lib()
    fn = L.sed_advsed1 if scheme else L.sed_advsed0
    fn.restype = None
    fn.argtypes = [C.c_int, _dp, _dp]
    c, y = _f8(c), _f8(y).copy()
    fn(len(y), _p(c), _p(y))
    return y


def sedp(g, dt, nf, detw, deta, t, p, vd, ff, diag):
    """str.f90:2257-2411 per column.  g: grid dict (nka, nkt, rq [nka,nkt], e [nkt], kw [nka]); t, p [ncol,n];
    vd [ncol,nka,nkt]; ff [ncol,n,nka,nkt]; diag [ncol,4] = (ajs, trdep, ds1, ds2).  Returns copies (ff, diag)."""
    L = _ko.lib()
    nka, nkt = int(g["nka"]), int(g["nkt"])
    detw, deta, t, p, vd = map(_f8, (detw, deta, t, p, vd))
    rq, e = _f8(g["rq"]), _f8(g["e"])
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    ff, diag = _f8(ff).copy(), _f8(diag).copy()
    ncol, n = t.shape
    assert ff.shape == (ncol, n, nka, nkt) and vd.shape == (ncol, nka, nkt) and diag.shape == (ncol, 4)
    L.sedp_oracle.restype = None
    L.sedp_oracle.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [_dp] * 6 + [
        C.POINTER(C.c_int32), _dp, _dp, _dp]
    L.sedp_oracle(ncol, n, int(nf), nka, nkt, float(dt), _p(detw), _p(deta), _p(t), _p(p), _p(rq), _p(e),
                  kw.ctypes.data_as(C.POINTER(C.c_int32)), _p(vd), _p(ff), _p(diag))
    return ff, diag


def sedl(dt, nf, nkc_l, detw, deta, t, p, rc, vt, vdm, s):
    """One half of str.f90:2627-2787 per column: s [ncol,n,nkc,jx] (sl1 or sion1); rc, vt [ncol,n,nkc];
    vdm [ncol,nkc].  Returns the updated copy."""
    L = _ko.lib()
    detw, deta, t, p, rc, vt, vdm = map(_f8, (detw, deta, t, p, rc, vt, vdm))
    s = _f8(s).copy()
    ncol, n, nkc, jx = s.shape
    assert rc.shape == (ncol, n, nkc) and vt.shape == (ncol, n, nkc) and vdm.shape == (ncol, nkc)
    L.sedl_oracle.restype = None
    L.sedl_oracle.argtypes = [C.c_int64] + [C.c_int] * 5 + [C.c_double] + [_dp] * 8
    L.sedl_oracle(ncol, n, int(nf), nkc, int(nkc_l), jx, float(dt), _p(detw), _p(deta), _p(t), _p(p), _p(rc), _p(vt),
                  _p(vdm), _p(s))
    return s


def sedc(dt, detw, deta, vg, es1, s1):
    """The species loop of str.f90:2567-2596 per column: s1 [ncol,n,j1]; vg, es1 [j1].  Returns the updated copy."""
    L = _ko.lib()
    detw, deta, vg, es1 = map(_f8, (detw, deta, vg, es1))
    s1 = _f8(s1).copy()
    ncol, n, j1 = s1.shape
    L.sedc_oracle.restype = None
    L.sedc_oracle.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_double] + [_dp] * 5
    L.sedc_oracle(ncol, n, j1, float(dt), _p(detw), _p(deta), _p(vg), _p(es1), _p(s1))
    return s1
