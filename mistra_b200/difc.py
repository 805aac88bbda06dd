"""Host-side mirror of SUBROUTINE difc (/root/reference/src/str.f90:3271-3445) over the C ABI of
include/mistra_difc.h: implicit turbulent exchange + subsidence of every chemical species of an ensemble
of columns, in place on the chemistry arrays.  CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kpp

MAXFIELDS = 8


class DifcField(C.Structure):
    _fields_ = [("s", C.c_void_p), ("row", C.c_int32), ("nproc", C.c_int32)]


class DifcArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("nfield", C.c_int32), ("dt", C.c_double)] + [
        (k, C.c_void_p) for k in ("atkh", "w", "am3", "detw", "deta")] + [("field", DifcField * MAXFIELDS)]


class DifpArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("row", C.c_int32), ("dt", C.c_double)] + [
        (k, C.c_void_p) for k in ("atkh", "w", "rho", "detw", "deta", "ff", "fsum")]


def _lib():
    L = kpp.library()
    L.mistra_difp.argtypes = [C.c_int64, C.POINTER(DifpArgs), C.c_void_p]
    L.mistra_difp_device.argtypes = [C.c_int64, C.POINTER(DifpArgs), C.c_void_p]
    L.mistra_difc.argtypes = [C.c_int64, C.POINTER(DifcArgs), C.c_void_p]
    L.mistra_difc_device.argtypes = [C.c_int64, C.POINTER(DifcArgs), C.c_void_p]
    L.mistra_difc_launch_count.restype = C.c_int64
    return L


def synthetic_columns(ncol, n=150, seed=0):
    """Columns for tests and bench: a stretched vertical grid (10 m layers below 1 km, growing above, as
    grid.f90 builds it), exchange coefficients of a boundary layer (0.01 - 50 m2/s), weak subsidence,
    air density falling with height."""
    r = np.random.default_rng(seed)
    detw = np.where(np.arange(n) < 100, 10.0, 10.0 * 1.1 ** (np.arange(n) - 99.0))
    deta = 0.5 * (detw + np.roll(detw, -1)); deta[-1] = detw[-1]
    z = np.cumsum(detw)
    zi = r.uniform(300.0, 1500.0, (ncol, 1))
    atkh = 0.01 + 50.0 * (z[None] / zi) * np.clip(1.0 - z[None] / zi, 0.0, None) ** 2 * r.uniform(0.5, 1.5, (ncol, n))
    w = -r.uniform(0.0, 0.01, (ncol, 1)) * np.minimum(z[None] / 1000.0, 1.0)
    w = np.where(r.uniform(size=(ncol, 1)) < 0.2, 0.0, w) * np.ones((1, n))
    am3 = 42.0 * np.exp(-z[None] / 8000.0) * r.uniform(0.97, 1.03, (ncol, n))
    return dict(detw=detw, deta=deta, atkh=atkh, w=w, am3=am3)


def difc(dt, atkh, w, am3, detw, deta, fields):
    """HOST numpy arrays: atkh, w, am3 [ncol,n]; detw, deta [n]; fields = list of (array [ncol,n,row],
    nproc): species 0..nproc-1 of each row are diffused.  Returns the updated copies."""
    L = _lib()
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    atkh, w, am3, detw, deta = map(f8, (atkh, w, am3, detw, deta))
    if atkh.ndim != 2:
        raise ValueError("difc: atkh must be [ncol,n]")
    ncol, n = atkh.shape
    if w.shape != (ncol, n) or am3.shape != (ncol, n) or detw.shape != (n,) or deta.shape != (n,):
        raise ValueError("difc: w, am3 [ncol,n]; detw, deta [n]")
    if len(fields) > MAXFIELDS:
        raise ValueError("difc: at most %d fields" % MAXFIELDS)
    outs = []
    a = DifcArgs(n, len(fields), float(dt), *[x.ctypes.data for x in (atkh, w, am3, detw, deta)])
    for i, (arr, nproc) in enumerate(fields):
        o = f8(arr).copy()
        if o.ndim != 3 or o.shape[:2] != (ncol, n):
            raise ValueError("difc: field %d must be [ncol,n,row]" % i)
        outs.append(o)
        a.field[i] = DifcField(o.ctypes.data, o.shape[2], int(nproc))
    kpp._check(L, L.mistra_difc(ncol, C.byref(a), None))
    return outs


def difc_device(dt, atkh, w, am3, detw, deta, fields, stream=None):
    """Same on torch CUDA tensors of the current device; the field tensors are updated in place.
    Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib()
    ncol, n = atkh.shape

    def ok(x, shape):
        if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float64 and tuple(x.shape) == shape):
            raise ValueError("difc_device: need contiguous CUDA float64 %s" % (shape,))
        return x.data_ptr()
    a = DifcArgs(n, len(fields), float(dt), ok(atkh, (ncol, n)), ok(w, (ncol, n)), ok(am3, (ncol, n)),
                 ok(detw, (n,)), ok(deta, (n,)))
    if len(fields) > MAXFIELDS:
        raise ValueError("difc_device: at most %d fields" % MAXFIELDS)
    for i, (arr, nproc) in enumerate(fields):
        if arr.dim() != 3:
            raise ValueError("difc_device: field %d must be [ncol,n,row]" % i)
        a.field[i] = DifcField(ok(arr, (ncol, n, arr.shape[2])), arr.shape[2], int(nproc))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_difc_device(ncol, C.byref(a), C.c_void_p(stream)))


def difp(dt, atkh, w, rho, detw, deta, ff, fsum):
    """SUBROUTINE difp (str.f90:3137-3265) on HOST numpy arrays: atkh, w, rho [ncol,n]; detw, deta [n];
    ff [ncol,n,nka*nkt] (or [ncol,n,nka,nkt]); fsum [ncol,n].  Returns updated copies (ff, fsum)."""
    L = _lib()
    f8 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    atkh, w, rho, detw, deta = map(f8, (atkh, w, rho, detw, deta))
    ff, fsum = f8(ff).copy(), f8(fsum).copy()
    if atkh.ndim != 2 or ff.ndim < 3:
        raise ValueError("difp: atkh [ncol,n], ff [ncol,n,row]")
    ncol, n = atkh.shape
    if ff.shape[:2] != (ncol, n) or fsum.shape != (ncol, n) or w.shape != (ncol, n) or rho.shape != (ncol, n) \
            or detw.shape != (n,) or deta.shape != (n,):
        raise ValueError("difp: w, rho, fsum [ncol,n]; detw, deta [n]; ff [ncol,n,row]")
    row = int(np.prod(ff.shape[2:]))
    a = DifpArgs(n, row, float(dt), *[x.ctypes.data for x in (atkh, w, rho, detw, deta, ff, fsum)])
    kpp._check(L, L.mistra_difp(ncol, C.byref(a), None))
    return ff, fsum


def difp_device(dt, atkh, w, rho, detw, deta, ff, fsum, stream=None):
    """Same on torch CUDA tensors of the current device; ff [ncol,n,row] and fsum are updated in place."""
    import torch
    L = _lib()
    ncol, n = atkh.shape

    def ok(x, shape):
        if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float64 and tuple(x.shape) == shape):
            raise ValueError("difp_device: need contiguous CUDA float64 %s" % (shape,))
        return x.data_ptr()
    if ff.dim() != 3:
        raise ValueError("difp_device: ff must be [ncol,n,row]")
    row = ff.shape[2]
    a = DifpArgs(n, row, float(dt), ok(atkh, (ncol, n)), ok(w, (ncol, n)), ok(rho, (ncol, n)), ok(detw, (n,)),
                 ok(deta, (n,)), ok(ff, (ncol, n, row)), ok(fsum, (ncol, n)))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_difp_device(ncol, C.byref(a), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_difc_launch_count())
