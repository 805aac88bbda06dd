"""Host-side mirror of SUBROUTINE konc (/root/reference/src/kpp.f90:3370-3585) over the C ABI of
include/mistra_konc.h: the dissolved species of the four chemistry bins follow the liquid volume
that the condensation step moved between the aerosol and the droplet part of every dry class.
Takes what `kon.kon_layers(..., chem=True)` leaves behind.  CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kpp

SUMS = ("vol1_a", "vol1_d", "part_o_a", "part_o_d", "part_n_a", "part_n_d")
J2, J6 = 121, 55          # global_params.f90: j2, j6


class KoncArgs(C.Structure):
    _fields_ = [("nka", C.c_int32), ("ka", C.c_int32), ("j2", C.c_int32), ("j6", C.c_int32)] + [
        (n, C.c_void_p) for n in SUMS + ("vol2", "pntot", "sl1", "sion1", "warn")]


def _lib():
    L = kpp.library()
    L.mistra_konc.argtypes = [C.c_int64, C.POINTER(KoncArgs), C.c_void_p]
    L.mistra_konc_device.argtypes = [C.c_int64, C.POINTER(KoncArgs), C.c_void_p]
    L.mistra_konc_launch_count.restype = C.c_int64
    return L


def konc(ka, sums, vol2, pntot, sl1, sion1):
    """HOST numpy arrays.  sums = dict of the six [ncell,nka] arrays of SUMS; vol2, pntot
    [ncell,4]; sl1 [ncell,4,j2], sion1 [ncell,4,j6] (not modified).
    Returns (sl1_new, sion1_new, warn[ncell,3])."""
    L = _lib()
    s = {k: np.ascontiguousarray(sums[k], dtype=np.float64) for k in SUMS}
    ncell, nka = s["vol1_a"].shape
    for k in SUMS:
        if s[k].shape != (ncell, nka):
            raise ValueError("konc: %s must be [ncell,nka]" % k)
    vol2 = np.ascontiguousarray(vol2, dtype=np.float64)
    pntot = np.ascontiguousarray(pntot, dtype=np.float64)
    if vol2.shape != (ncell, 4) or pntot.shape != (ncell, 4):
        raise ValueError("konc: vol2, pntot must be [ncell,4]")
    sl1 = np.array(sl1, dtype=np.float64, order="C")
    sion1 = np.array(sion1, dtype=np.float64, order="C")
    if sl1.ndim != 3 or sion1.ndim != 3 or sl1.shape[:2] != (ncell, 4) or sion1.shape[:2] != (ncell, 4):
        raise ValueError("konc: sl1, sion1 must be [ncell,4,j]")
    warn = np.zeros((ncell, 3), dtype=np.int32)
    a = KoncArgs(nka, int(ka), sl1.shape[2], sion1.shape[2], *[s[k].ctypes.data for k in SUMS],
                 vol2.ctypes.data, pntot.ctypes.data, sl1.ctypes.data, sion1.ctypes.data, warn.ctypes.data)
    kpp._check(L, L.mistra_konc(ncell, C.byref(a), None))
    return sl1, sion1, warn


def konc_device(ka, sums, vol2, pntot, sl1, sion1, warn=None, stream=None):
    """Same on torch CUDA tensors of the current device; sl1, sion1 are updated in place.
    Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib()
    ncell, nka = sums["vol1_a"].shape

    def ok(t, shape, dt=torch.float64):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and tuple(t.shape) == shape):
            raise ValueError("konc_device: need contiguous CUDA %s %s" % (dt, shape))
        return t.data_ptr()
    ptrs = [ok(sums[k], (ncell, nka)) for k in SUMS]
    j2, j6 = sl1.shape[2], sion1.shape[2]
    a = KoncArgs(nka, int(ka), j2, j6, *ptrs, ok(vol2, (ncell, 4)), ok(pntot, (ncell, 4)),
                 ok(sl1, (ncell, 4, j2)), ok(sion1, (ncell, 4, j6)),
                 ok(warn, (ncell, 3), torch.int32) if warn is not None else None)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_konc_device(ncell, C.byref(a), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_konc_launch_count())


def synthetic_sums(ncell, nka=70, ka=32, seed=20261018, j2=J2, j6=J6):
    """Synthetic inputs of konc shaped like kon's bin sums: per dry class a particle number
    split between aerosol and droplet part before the step and a (mostly small) number of
    particles that changed side; a few classes violate number conservation or come out with
    delta outside [0, 1] (the reference's warning paths), some layers have almost no droplets
    left (the transfer of kpp.f90:3566-3586), some bins are empty."""
    rng = np.random.default_rng(seed)
    tot = 10.0 ** rng.uniform(-3, 3, (ncell, nka))
    fa = rng.uniform(0, 1, (ncell, nka))
    fa[rng.uniform(size=fa.shape) < 0.2] = 1.0              # class entirely on the aerosol side
    fa[rng.uniform(size=fa.shape) < 0.05] = 0.0             # ... or on the droplet side
    part_o_a, part_o_d = tot * fa, tot * (1 - fa)
    move = rng.uniform(-0.3, 0.3, (ncell, nka)) * (rng.uniform(size=(ncell, nka)) < 0.5)
    x = np.where(move > 0, move * part_o_a, move * part_o_d)   # > 0: aerosol -> droplets
    part_n_a, part_n_d = part_o_a - x, part_o_d + x
    bad = rng.uniform(size=x.shape) < 0.01
    part_n_d = np.where(bad, part_n_d * 1.001 + 1e-6, part_n_d)
    worse = rng.uniform(size=x.shape) < 0.01                   # droplets appear from nowhere: delta < 0
    part_n_d = np.where(worse, part_n_d * 2.0 + 1e-3, part_n_d)
    r3a, r3d = 10.0 ** rng.uniform(-3, 0, (ncell, nka)), 10.0 ** rng.uniform(0, 3, (ncell, nka))
    vol1_a, vol1_d = part_o_a * r3a, part_o_d * r3d
    vol2 = np.zeros((ncell, 4))
    vol2[:, 0] = vol1_a[:, :ka].sum(1); vol2[:, 1] = vol1_a[:, ka:].sum(1)
    vol2[:, 2] = vol1_d[:, :ka].sum(1); vol2[:, 3] = vol1_d[:, ka:].sum(1)
    odd = rng.uniform(size=ncell) < 0.05
    vol2[odd] *= 10.0 ** rng.uniform(-4, 0, (int(odd.sum()), 4))   # inconsistent sums: delta > 1 shows up
    vol2[rng.uniform(size=ncell) < 0.03, 2] = 0.0
    pntot = np.zeros((ncell, 4))
    pntot[:, 0] = part_n_a[:, :ka].sum(1); pntot[:, 1] = part_n_a[:, ka:].sum(1)
    pntot[:, 2] = part_n_d[:, :ka].sum(1); pntot[:, 3] = part_n_d[:, ka:].sum(1)
    few = rng.uniform(size=(ncell, 2)) < 0.1
    pntot[:, 2:] = np.where(few, 1e-9, pntot[:, 2:])
    sl1 = 10.0 ** rng.uniform(-14, -6, (ncell, 4, j2)) * (rng.uniform(size=(ncell, 4, j2)) < 0.7)
    sion1 = 10.0 ** rng.uniform(-12, -5, (ncell, 4, j6)) * (rng.uniform(size=(ncell, 4, j6)) < 0.8)
    sl1[rng.uniform(size=sl1.shape) < 0.01] *= -1.0            # slightly negative leftovers get clipped
    sums = dict(vol1_a=vol1_a, vol1_d=vol1_d, part_o_a=part_o_a, part_o_d=part_o_d,
                part_n_a=part_n_a, part_n_d=part_n_d)
    return dict(ka=ka, sums=sums, vol2=vol2, pntot=pntot, sl1=sl1, sion1=sion1)
