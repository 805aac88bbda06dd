"""Host-side mirror of SUBROUTINE cw_rc (/root/reference/src/kpp.f90:2152-2414) over the C ABI of
include/mistra_cwrc.h: liquid water content cw, mean radius rc, water volume cm and the switch /
conversion factor conv2 of the four chemistry bins from the particle spectrum ff.
CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kpp

CRYS = dict(xcryssulf=0.4, xcrysss=0.42, xdelisulf=0.7, xdeliss=0.75)     # initc, kpp.f90:319-324


class CwrcArgs(C.Structure):
    _fields_ = [("nka", C.c_int32), ("nkt", C.c_int32), ("ka", C.c_int32), ("ial", C.c_int32),
                ("xcryssulf", C.c_double), ("xcrysss", C.c_double), ("xdelisulf", C.c_double),
                ("xdeliss", C.c_double)] + [
        (n, C.c_void_p) for n in ("kw", "e", "rq", "ff", "feu", "cloud", "rc", "cw", "cm", "conv2")]


def _lib():
    L = kpp.library()
    L.mistra_cwrc.argtypes = [C.c_int64, C.POINTER(CwrcArgs), C.c_void_p]
    L.mistra_cwrc_device.argtypes = [C.c_int64, C.POINTER(CwrcArgs), C.c_void_p]
    L.mistra_cwrc_launch_count.restype = C.c_int64
    return L


def cw_rc(g, ff, feu, cloud, ial=1, crys=None):
    """HOST numpy arrays: g = grid dict with nka, nkt, ka, kw, e, rq (kon.kon_grid()); ff
    [ncell,nka,nkt], feu [ncell], cloud [ncell,4] (bool / int).  Returns (rc, cw, cm, conv2),
    each [ncell,4]."""
    L = _lib()
    cr = dict(CRYS, **(crys or {}))
    nka, nkt = int(g["nka"]), int(g["nkt"])
    ff = np.ascontiguousarray(ff, dtype=np.float64)
    n = ff.shape[0]
    if ff.shape != (n, nka, nkt):
        raise ValueError("cw_rc: ff must be [ncell,nka,nkt]")
    feu = np.ascontiguousarray(feu, dtype=np.float64)
    cloud = np.ascontiguousarray(cloud).astype(np.int32)
    if feu.shape != (n,) or cloud.shape != (n, 4):
        raise ValueError("cw_rc: feu [ncell], cloud [ncell,4]")
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    e = np.ascontiguousarray(g["e"], dtype=np.float64)
    rq = np.ascontiguousarray(g["rq"], dtype=np.float64)
    out = [np.zeros((n, 4)) for _ in range(4)]
    a = CwrcArgs(nka, nkt, int(g["ka"]), int(ial), cr["xcryssulf"], cr["xcrysss"], cr["xdelisulf"], cr["xdeliss"],
                 kw.ctypes.data, e.ctypes.data, rq.ctypes.data, ff.ctypes.data, feu.ctypes.data, cloud.ctypes.data,
                 *[o.ctypes.data for o in out])
    kpp._check(L, L.mistra_cwrc(n, C.byref(a), None))
    return tuple(out)


def cw_rc_device(g_dev, ff, feu, cloud, rc, cw, cm, conv2, ial=1, crys=None, stream=None):
    """Same on torch CUDA tensors of the current device.  g_dev: dict with nka, nkt, ka and CUDA
    tensors kw (int32 [nka]), e ([nkt]), rq ([nka,nkt]); outputs are written in place.
    Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib()
    cr = dict(CRYS, **(crys or {}))
    nka, nkt = int(g_dev["nka"]), int(g_dev["nkt"])
    n = ff.shape[0]

    def ok(t, shape, dt=torch.float64):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and tuple(t.shape) == shape):
            raise ValueError("cw_rc_device: need contiguous CUDA %s %s" % (dt, shape))
        return t.data_ptr()
    a = CwrcArgs(nka, nkt, int(g_dev["ka"]), int(ial), cr["xcryssulf"], cr["xcrysss"], cr["xdelisulf"], cr["xdeliss"],
                 ok(g_dev["kw"], (nka,), torch.int32), ok(g_dev["e"], (nkt,)), ok(g_dev["rq"], (nka, nkt)),
                 ok(ff, (n, nka, nkt)), ok(feu, (n,)), ok(cloud, (n, 4), torch.int32),
                 ok(rc, (n, 4)), ok(cw, (n, 4)), ok(cm, (n, 4)), ok(conv2, (n, 4)))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_cwrc_device(n, C.byref(a), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_cwrc_launch_count())
