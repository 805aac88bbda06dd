"""Host-side mirror of SUBROUTINE sedp, sedl and sedc (/root/reference/src/str.f90:2257-2411, 2627-2787, 2567-2596)
over the C ABI of include/mistra_sed.h: gravitational settling of the particle spectrum and of the aqueous species,
dry deposition / emission of the gases, for an ensemble of columns, in place on the chemistry arrays.
CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kpp


class SedpArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("nf", C.c_int32), ("nka", C.c_int32), ("nkt", C.c_int32), ("dt", C.c_double)] + [
        (k, C.c_void_p) for k in ("detw", "deta", "t", "p", "rq", "e", "kw", "vd", "ff", "diag")]


class SedlArgs(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "nf", "nkc", "nkc_l", "j2", "j6")] + [("dt", C.c_double)] + [
        (k, C.c_void_p) for k in ("detw", "deta", "t", "p", "rc", "vt", "vdm", "sl1", "sion1")]


class SedcArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("j1", C.c_int32), ("dt", C.c_double)] + [
        (k, C.c_void_p) for k in ("detw", "deta", "vg", "es1", "s1")]


def _lib():
    L = kpp.library()
    for name, cls in (("sedp", SedpArgs), ("sedl", SedlArgs), ("sedc", SedcArgs)):
        getattr(L, "mistra_%s" % name).argtypes = [C.c_int64, C.POINTER(cls), C.c_void_p]
        getattr(L, "mistra_%s_device" % name).argtypes = [C.c_int64, C.POINTER(cls), C.c_void_p]
    L.mistra_sed_launch_count.restype = C.c_int64
    L.mistra_sed_divc_selftest.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    return L


_f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)


def synthetic_columns(g, ncol, n=150, nf=100, seed=0, nkc=4, j2=121, j6=55, populated=0.06):
    """Columns for tests and bench: the vertical grid of difc.synthetic_columns, a standard atmosphere, a particle
    spectrum that populates `populated` of the 2-D grid (a marine aerosol band plus, in a third of the columns, a
    cloud layer with drops up to drizzle size), bin radii / velocities of the aqueous bins as fast_k_mt leaves them."""
    from . import difc
    r = np.random.default_rng(seed)
    c = difc.synthetic_columns(ncol, n, seed)
    detw, deta = c["detw"], c["deta"]
    z = np.cumsum(detw)
    t = 288.0 - 0.0065 * z[None] + r.uniform(-1.0, 1.0, (ncol, n))
    p = 101325.0 * np.exp(-z[None] / 8000.0) * np.ones((ncol, 1))
    nka, nkt = int(g["nka"]), int(g["nkt"])
    rq = _f8(g["rq"])
    ff = np.zeros((ncol, n, nka, nkt))
    band = r.uniform(size=(nka, nkt)) < populated * 2.0
    band &= rq < 2.0
    prof = np.exp(-z / 1200.0)[None, :, None, None]
    ff += band[None, None] * prof * 10.0 ** r.uniform(-1, 2.5, (ncol, 1, nka, nkt))
    cloudy = r.uniform(size=ncol) < 1.0 / 3.0
    kb, kt = r.integers(20, 40, ncol), r.integers(45, 70, ncol)
    drops = (r.uniform(size=(nka, nkt)) < populated) & (rq > 3.0) & (rq < 300.0)
    for col in np.nonzero(cloudy)[0]:
        ff[col, kb[col]:min(kt[col], nf - 2)] += drops[None] * 10.0 ** r.uniform(-4, 1, (1, nka, nkt))
    vd = 10.0 ** r.uniform(-4, -1.5, (ncol, nka, nkt))
    diag = np.zeros((ncol, 4)); diag[:, 1:] = r.uniform(0, 1e-3, (ncol, 3))
    rc = np.zeros((ncol, n, nkc)); vt = np.zeros((ncol, n, nkc))
    rc[..., 0], rc[..., 1] = 10.0 ** r.uniform(-7.5, -6.5, (ncol, n)), 10.0 ** r.uniform(-6.3, -5.5, (ncol, n))
    rc[..., 2:] = np.where(cloudy[:, None, None], 10.0 ** r.uniform(-5.3, -4.3, (ncol, n, nkc - 2)), 0.0)
    vt[:, :nf] = 1.2e8 * rc[:, :nf] ** 2 * r.uniform(0.8, 1.5, (ncol, nf, nkc))
    vdm = 10.0 ** r.uniform(-4, -2, (ncol, nkc))
    sl1 = 10.0 ** r.uniform(-13, -7, (ncol, n, nkc, j2)) * (r.uniform(size=(ncol, n, nkc, j2)) < 0.7)
    sion1 = 10.0 ** r.uniform(-13, -7, (ncol, n, nkc, j6)) * (r.uniform(size=(ncol, n, nkc, j6)) < 0.7)
    return dict(detw=detw, deta=deta, t=t, p=p, ff=ff, vd=vd, diag=diag, rc=rc, vt=vt, vdm=vdm, sl1=sl1, sion1=sion1)


def sedp(g, dt, nf, detw, deta, t, p, vd, ff, diag):
    """HOST numpy arrays.  g: grid dict with nka, nkt, rq [nka,nkt], e [nkt], kw [nka] (kon.kon_grid()); t, p
    [ncol,n]; vd [ncol,nka,nkt]; ff [ncol,n,nka,nkt]; diag [ncol,4] = (ajs, trdep, ds1, ds2).  Returns updated
    copies (ff, diag)."""
    L = _lib()
    nka, nkt = int(g["nka"]), int(g["nkt"])
    detw, deta, t, p, vd = map(_f8, (detw, deta, t, p, vd))
    rq, e = _f8(g["rq"]), _f8(g["e"])
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    ff, diag = _f8(ff).copy(), _f8(diag).copy()
    if t.ndim != 2:
        raise ValueError("sedp: t must be [ncol,n]")
    ncol, n = t.shape
    if ff.shape != (ncol, n, nka, nkt) or vd.shape != (ncol, nka, nkt) or diag.shape != (ncol, 4) or p.shape != (ncol, n) \
            or detw.shape != (n,) or deta.shape != (n,) or rq.shape != (nka, nkt) or e.shape != (nkt,) or kw.shape != (nka,):
        raise ValueError("sedp: ff [ncol,n,nka,nkt]; vd [ncol,nka,nkt]; diag [ncol,4]; t, p [ncol,n]; detw, deta [n]")
    a = SedpArgs(n, int(nf), nka, nkt, float(dt), *[x.ctypes.data for x in (detw, deta, t, p, rq, e, kw, vd, ff, diag)])
    kpp._check(L, L.mistra_sedp(ncol, C.byref(a), None))
    return ff, diag


def sedl(dt, nf, nkc_l, detw, deta, t, p, rc, vt, vdm, sl1=None, sion1=None):
    """HOST numpy arrays: rc, vt [ncol,n,nkc]; vdm [ncol,nkc]; sl1 [ncol,n,nkc,j2], sion1 [ncol,n,nkc,j6] (either
    may be None).  Returns updated copies (sl1, sion1)."""
    L = _lib()
    detw, deta, t, p, rc, vt, vdm = map(_f8, (detw, deta, t, p, rc, vt, vdm))
    if rc.ndim != 3:
        raise ValueError("sedl: rc must be [ncol,n,nkc]")
    ncol, n, nkc = rc.shape
    if vt.shape != (ncol, n, nkc) or vdm.shape != (ncol, nkc) or t.shape != (ncol, n) or p.shape != (ncol, n) \
            or detw.shape != (n,) or deta.shape != (n,):
        raise ValueError("sedl: rc, vt [ncol,n,nkc]; vdm [ncol,nkc]; t, p [ncol,n]; detw, deta [n]")
    outs, ptr, jx = [], [], []
    for s in (sl1, sion1):
        if s is None:
            outs.append(None); ptr.append(None); jx.append(0)
            continue
        o = _f8(s).copy()
        if o.ndim != 4 or o.shape[:3] != (ncol, n, nkc):
            raise ValueError("sedl: sl1 / sion1 must be [ncol,n,nkc,j]")
        outs.append(o); ptr.append(o.ctypes.data); jx.append(o.shape[3])
    a = SedlArgs(n, int(nf), nkc, int(nkc_l), jx[0], jx[1], float(dt),
                 *[x.ctypes.data for x in (detw, deta, t, p, rc, vt, vdm)], ptr[0], ptr[1])
    kpp._check(L, L.mistra_sedl(ncol, C.byref(a), None))
    return tuple(outs)


def sedc(dt, detw, deta, vg, es1, s1):
    """HOST numpy arrays: s1 [ncol,n,j1]; vg, es1 [j1].  Returns the updated copy."""
    L = _lib()
    detw, deta, vg, es1 = map(_f8, (detw, deta, vg, es1))
    s1 = _f8(s1).copy()
    if s1.ndim != 3:
        raise ValueError("sedc: s1 must be [ncol,n,j1]")
    ncol, n, j1 = s1.shape
    if vg.shape != (j1,) or es1.shape != (j1,) or detw.shape != (n,) or deta.shape != (n,):
        raise ValueError("sedc: vg, es1 [j1]; detw, deta [n]")
    a = SedcArgs(n, j1, float(dt), *[x.ctypes.data for x in (detw, deta, vg, es1, s1)])
    kpp._check(L, L.mistra_sedc(ncol, C.byref(a), None))
    return s1


def _ok(torch, name):
    def ok(x, shape, dt=None):
        dt = dt or torch.float64
        if not (x.is_cuda and x.is_contiguous() and x.dtype == dt and tuple(x.shape) == tuple(shape)):
            raise ValueError("%s: need contiguous CUDA %s %s" % (name, dt, tuple(shape)))
        return x.data_ptr()
    return ok


def sedp_device(g_dev, dt, nf, detw, deta, t, p, vd, ff, diag, stream=None):
    """Same on torch CUDA tensors of the current device (g_dev: nka, nkt and CUDA tensors rq, e, kw int32); ff and
    diag are updated in place.  Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib()
    ok = _ok(torch, "sedp_device")
    nka, nkt = int(g_dev["nka"]), int(g_dev["nkt"])
    ncol, n = t.shape
    a = SedpArgs(n, int(nf), nka, nkt, float(dt), ok(detw, (n,)), ok(deta, (n,)), ok(t, (ncol, n)), ok(p, (ncol, n)),
                 ok(g_dev["rq"], (nka, nkt)), ok(g_dev["e"], (nkt,)), ok(g_dev["kw"], (nka,), torch.int32),
                 ok(vd, (ncol, nka, nkt)), ok(ff, (ncol, n, nka, nkt)), ok(diag, (ncol, 4)))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_sedp_device(ncol, C.byref(a), C.c_void_p(stream)))


def sedl_device(dt, nf, nkc_l, detw, deta, t, p, rc, vt, vdm, sl1=None, sion1=None, stream=None):
    """Same on torch CUDA tensors; sl1 / sion1 are updated in place."""
    import torch
    L = _lib()
    ok = _ok(torch, "sedl_device")
    ncol, n, nkc = rc.shape
    j2 = sl1.shape[3] if sl1 is not None else 0
    j6 = sion1.shape[3] if sion1 is not None else 0
    a = SedlArgs(n, int(nf), nkc, int(nkc_l), j2, j6, float(dt), ok(detw, (n,)), ok(deta, (n,)), ok(t, (ncol, n)),
                 ok(p, (ncol, n)), ok(rc, (ncol, n, nkc)), ok(vt, (ncol, n, nkc)), ok(vdm, (ncol, nkc)),
                 ok(sl1, (ncol, n, nkc, j2)) if sl1 is not None else None,
                 ok(sion1, (ncol, n, nkc, j6)) if sion1 is not None else None)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_sedl_device(ncol, C.byref(a), C.c_void_p(stream)))


def sedc_device(dt, detw, deta, vg, es1, s1, stream=None):
    """Same on torch CUDA tensors; s1 [ncol,n,j1] is updated in place."""
    import torch
    L = _lib()
    ok = _ok(torch, "sedc_device")
    ncol, n, j1 = s1.shape
    a = SedcArgs(n, j1, float(dt), ok(detw, (n,)), ok(deta, (n,)), ok(vg, (j1,)), ok(es1, (j1,)), ok(s1, (ncol, n, j1)))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_sedc_device(ncol, C.byref(a), C.c_void_p(stream)))


def divc_selftest(n, seed=1):
    """Number of arguments (of n) for which advsed1's multiply-and-correct divisions by constants differ from the IEEE
    division on the device (expected 0)."""
    L = _lib()
    bad = C.c_int64(-1)
    kpp._check(L, L.mistra_sed_divc_selftest(int(n), int(seed), C.byref(bad)))
    return int(bad.value)


def launch_count():
    return int(_lib().mistra_sed_launch_count())
