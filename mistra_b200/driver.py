"""Host-side mirror of the layer loop of SUBROUTINE kpp_driver (/root/reference/src/kpp.f90:4305-4470) over the C ABI
of include/mistra_driver.h: the per-layer scalars and switches of every *_drive call, the cloud bookkeeping, the
Eulerian advection source, and the layers sorted by mechanism.  CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kpp

NPHRXN = 47


class DriverArgs(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "nf", "nkc", "nphrxn", "halo", "iod", "lpBuys13_0D", "neula", "box", "n_bl",
                                         "kinv", "nadv", "j1", "j5")] + [("dt_ch", C.c_double)] + [
        (k, C.c_void_p) for k in ("u0", "t", "p", "rho", "cm3", "am3", "xm1", "conv2", "cm", "cloud", "photol_j", "adv_row",
                                  "xadv", "s1", "s3", "cb1", "scal", "ph_rat", "air", "h2o", "cvv", "mech", "layers", "count")]


def _lib():
    L = kpp.library()
    L.mistra_driver_layers.argtypes = [C.c_int64, C.POINTER(DriverArgs), C.c_void_p]
    L.mistra_driver_layers_device.argtypes = [C.c_int64, C.POINTER(DriverArgs), C.c_void_p]
    L.mistra_driver_launch_count.restype = C.c_int64
    return L


def _args(cfg, n, nkc, nph, nadv, j1, j5, ptr):
    return DriverArgs(n, int(cfg["nf"]), nkc, nph, int(bool(cfg["halo"])), int(bool(cfg["iod"])), int(bool(cfg["lpBuys13_0D"])),
                      int(cfg["neula"]), int(bool(cfg["box"])), int(cfg["n_bl"]), int(cfg["kinv"]), nadv, j1, j5,
                      float(cfg["dt_ch"]), *ptr)


def synthetic_columns(ncol, n, nf, nkc, j1, j5, seed):
    """Columns for tests, fixtures and bench: a standard atmosphere, water volume in the aerosol bins of most
    layers and in the droplet bins of some columns, a few night columns, a few negative entries in s1 / s3."""
    r = np.random.default_rng(seed)
    z = np.cumsum(np.full(n, 10.0))
    d = dict(t=288.0 - 0.0065 * z[None] + r.uniform(-1, 1, (ncol, n)), p=101325.0 * np.exp(-z[None] / 8000.0) * np.ones((ncol, 1)))
    d["rho"] = d["p"] / (287.05 * d["t"])
    d["am3"] = d["p"] / (8.3144743 * d["t"])
    d["cm3"] = d["am3"] * 6.022140857e+17
    d["xm1"] = 10.0 ** r.uniform(-4, -2, (ncol, n))
    d["conv2"] = 10.0 ** r.uniform(6, 12, (ncol, n, nkc))
    cm = 10.0 ** r.uniform(-14, -9, (ncol, n, nkc))
    cm[..., 2:] *= (r.uniform(size=(ncol, 1, 1)) < 0.5) * (r.uniform(size=(ncol, n, nkc - 2)) < 0.6)   # cloud in some columns
    cm[..., :2] *= r.uniform(size=(ncol, n, 2)) < 0.9                                                    # dry layers
    d["cm"] = cm
    d["conv2"] = np.where(cm > 0, d["conv2"], 0.0)
    d["cloud"] = (r.uniform(size=(ncol, n, nkc)) < 0.5).astype(np.int32)
    d["photol_j"] = 10.0 ** r.uniform(-8, -2, (ncol, n, NPHRXN))
    d["u0"] = np.array([0.5, 0.02, 0.0348, 0.0347, -0.3, 0.9][:ncol] + [0.7] * max(0, ncol - 6))
    d["s1"] = 10.0 ** r.uniform(-12, -7, (ncol, n, j1)) * np.where(r.uniform(size=(ncol, n, j1)) < 0.05, -1.0, 1.0)
    d["s3"] = 10.0 ** r.uniform(-14, -10, (ncol, n, j5)) * np.where(r.uniform(size=(ncol, n, j5)) < 0.05, -1.0, 1.0)
    return d


def layers(cfg, u0, t, p, rho, cm3, am3, xm1, conv2, cm, cloud, photol_j, adv_row=None, xadv=None, s1=None, s3=None):
    """HOST numpy arrays (shapes as in include/mistra_driver.h).  cfg: dict(nf, halo, iod, lpBuys13_0D, neula, box,
    n_bl, kinv, dt_ch).  Returns a dict: cb1, scal, ph_rat, air, h2o, cvv, mech (rows L = col * n + k - 1; rows of
    levels outside n_min .. n_max zero, mech -1), layers (three ascending index arrays), cloud, s1, s3 (updated copies)."""
    L = _lib()
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    u0, t, p, rho, cm3, am3, xm1, conv2, cm, photol_j = map(f8, (u0, t, p, rho, cm3, am3, xm1, conv2, cm, photol_j))
    if t.ndim != 2 or cm.ndim != 3 or photol_j.ndim != 3:
        raise ValueError("driver.layers: t [ncol,n], cm [ncol,n,nkc], photol_j [ncol,n,nphrxn]")
    ncol, n = t.shape
    nkc, nph = cm.shape[2], photol_j.shape[2]
    for a, sh in ((u0, (ncol,)), (p, (ncol, n)), (rho, (ncol, n)), (cm3, (ncol, n)), (am3, (ncol, n)), (xm1, (ncol, n)),
                  (conv2, (ncol, n, nkc)), (photol_j, (ncol, n, nph))):
        if a.shape != sh:
            raise ValueError("driver.layers: bad shape %s, expected %s" % (a.shape, sh))
    cloud = np.ascontiguousarray(cloud).astype(np.int32)
    if cloud.shape != (ncol, n, nkc):
        raise ValueError("driver.layers: cloud [ncol,n,nkc]")
    s1 = None if s1 is None else f8(s1).copy()
    s3 = None if s3 is None else f8(s3).copy()
    adv_row = np.zeros(0, dtype=np.int32) if adv_row is None else np.ascontiguousarray(adv_row, dtype=np.int32)
    xadv = np.zeros(0) if xadv is None else f8(xadv)
    nl = ncol * n
    out = dict(cb1=np.zeros((nl, 4)), scal=np.zeros((nl, 13)), ph_rat=np.zeros((nl, nph)), air=np.zeros(nl), h2o=np.zeros(nl),
               cvv=np.zeros((nl, 4)), mech=np.full(nl, -1, dtype=np.int32))
    lay, count = np.zeros((3, max(nl, 1)), dtype=np.int64), np.zeros(3, dtype=np.int64)
    ptr = [x.ctypes.data for x in (u0, t, p, rho, cm3, am3, xm1, conv2, cm, cloud, photol_j)]
    ptr += [adv_row.ctypes.data if len(adv_row) else None, xadv.ctypes.data if len(xadv) else None]
    ptr += [None if s1 is None else s1.ctypes.data, None if s3 is None else s3.ctypes.data]
    ptr += [out[k].ctypes.data for k in ("cb1", "scal", "ph_rat", "air", "h2o", "cvv", "mech")] + [lay.ctypes.data, count.ctypes.data]
    a = _args(cfg, n, nkc, nph, len(adv_row), 0 if s1 is None else s1.shape[2], 0 if s3 is None else s3.shape[2], ptr)
    kpp._check(L, L.mistra_driver_layers(ncol, C.byref(a), None))
    out.update(layers=[lay[m, :count[m]].copy() for m in range(3)], cloud=cloud, s1=s1, s3=s3)
    return out


def layers_device(cfg, u0, t, p, rho, cm3, am3, xm1, conv2, cm, cloud, photol_j, out, adv_row=None, xadv=None, s1=None,
                  s3=None, stream=None):
    """Same on torch CUDA tensors of the current device.  `out`: dict of preallocated tensors cb1 [nl,4], scal [nl,13],
    ph_rat [nl,nphrxn], air, h2o [nl], cvv [nl,4], mech int32 [nl], layers int64 [3,nl], count int64 [3]; cloud (int32),
    s1, s3 are updated in place.  Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib()
    ncol, n = t.shape
    nkc, nph = cm.shape[2], photol_j.shape[2]
    nl = ncol * n

    def ok(x, shape, dt=torch.float64):
        if x is None:
            return None
        if not (x.is_cuda and x.is_contiguous() and x.dtype == dt and tuple(x.shape) == tuple(shape)):
            raise ValueError("driver.layers_device: need contiguous CUDA %s %s" % (dt, tuple(shape)))
        return x.data_ptr()
    nadv = 0 if adv_row is None else adv_row.shape[0]
    ptr = [ok(u0, (ncol,))] + [ok(x, (ncol, n)) for x in (t, p, rho, cm3, am3, xm1)] + [
        ok(conv2, (ncol, n, nkc)), ok(cm, (ncol, n, nkc)), ok(cloud, (ncol, n, nkc), torch.int32), ok(photol_j, (ncol, n, nph)),
        ok(adv_row, (nadv,), torch.int32), ok(xadv, (nadv,)),
        ok(s1, (ncol, n, s1.shape[2])) if s1 is not None else None, ok(s3, (ncol, n, s3.shape[2])) if s3 is not None else None,
        ok(out["cb1"], (nl, 4)), ok(out["scal"], (nl, 13)), ok(out["ph_rat"], (nl, nph)), ok(out["air"], (nl,)),
        ok(out["h2o"], (nl,)), ok(out["cvv"], (nl, 4)), ok(out["mech"], (nl,), torch.int32),
        ok(out["layers"], (3, nl), torch.int64), ok(out["count"], (3,), torch.int64)]
    a = _args(cfg, n, nkc, nph, nadv, 0 if s1 is None else s1.shape[2], 0 if s3 is None else s3.shape[2], ptr)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_driver_layers_device(ncol, C.byref(a), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_driver_launch_count())
