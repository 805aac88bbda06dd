"""Host-side mirror of the reference's KPP driver interface over the C ABI
(include/mistra_kpp.h, libmistra_kpp.so).

Names follow the reference: `integrate` is INTEGRATE_x(TIN,TOUT) for a batch of
cells (/root/reference/src/gas.f:710, aer.f:1408, tot.f:2812); `MECH_*` are the
three KPP mechanisms kpp_driver dispatches to (kpp.f90:4451-4468); the error codes
are those of ros_ErrorMsg_x (gas.f:1474-1509).

There is no CPU implementation behind this module: if the CUDA library is missing
or no device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MECH_GAS, MECH_AER, MECH_TOT = 0, 1, 2
MECH_NAMES = ("gas", "aer", "tot")
STAT_NAMES = ("Nfun", "Njac", "Nstp", "Nacc", "Nrej", "Ndec", "Nsol", "Nsng")


class KppError(RuntimeError):
    pass


class KppOpts(C.Structure):
    """RPAR/IPAR of Rosenbrock_x (gas.f:786-870); zero selects the reference default."""
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double),
                ("hmin", C.c_double), ("hmax", C.c_double), ("hstart", C.c_double),
                ("facmin", C.c_double), ("facmax", C.c_double), ("facrej", C.c_double),
                ("facsafe", C.c_double),
                ("max_steps", C.c_int32), ("autonomous", C.c_int32),
                ("f32_literals", C.c_int32), ("reserved", C.c_int32)]


_LIBS = {}


def library(strict=False):
    """Load libmistra_kpp.so (or the -DKPP_STRICT -fmad=false test build)."""
    name = "libmistra_kpp_strict.so" if strict else os.environ.get("MISTRA_KPP_LIB", "libmistra_kpp.so")
    if name not in _LIBS:
        so = os.path.join(_HERE, name)
        if not os.path.exists(so):
            raise KppError("%s is missing: build it with `python -m mistra_b200.build` "
                           "(there is no CPU fallback)" % so)
        L = C.CDLL(so)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
        L.mistra_kpp_default_opts.argtypes = [C.POINTER(KppOpts)]
        L.mistra_kpp_default_opts.restype = None
        L.mistra_kpp_query.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4
        L.mistra_kpp_spc_name.argtypes = [C.c_int, C.c_int]
        L.mistra_kpp_spc_name.restype = C.c_char_p
        L.mistra_kpp_integrate.argtypes = [C.c_int, C.c_int64, dp, dp, dp, C.c_double, C.c_double,
                                           C.POINTER(KppOpts), ip, ip, dp, dp, vp]
        L.mistra_kpp_integrate_device.argtypes = [C.c_int, C.c_int64, vp, vp, vp, C.c_double,
                                                  C.c_double, C.POINTER(KppOpts), vp, vp, vp, vp, vp]
        L.mistra_kpp_integrate_multi.argtypes = [C.c_int, C.c_int64, dp, dp, dp, C.c_double, C.c_double,
                                                 C.POINTER(KppOpts), ip, ip, dp, dp, C.c_int, C.POINTER(C.c_int)]
        L.mistra_kpp_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.mistra_kpp_host_free.argtypes = [C.c_void_p]
        L.mistra_kpp_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.mistra_kpp_host_unregister.argtypes = [C.c_void_p]
        L.mistra_kpp_launch_count.restype = C.c_int64
        L.mistra_kpp_fp64_peak_tflops.restype = C.c_double
        L.mistra_kpp_last_error.restype = C.c_char_p
        _LIBS[name] = L
    return _LIBS[name]


def _check(L, rc):
    if rc != 0:
        raise KppError("mistra_kpp error %d: %s" % (rc, (L.mistra_kpp_last_error() or b"").decode()))


def default_opts(strict=False, **kw):
    o = KppOpts()
    library(strict).mistra_kpp_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def query(mech, strict=False):
    """(NVAR, NFIX, NREACT, LU_NONZERO) of a mechanism (*_Parameters.h)."""
    L = library(strict)
    v = [C.c_int() for _ in range(4)]
    _check(L, L.mistra_kpp_query(mech, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


def spc_name(mech, i, strict=False):
    s = library(strict).mistra_kpp_spc_name(mech, i)
    return s.decode() if s else None


def integrate(mech, rconst, fix, var, t0=0.0, t1=10.0, opts=None, strict=False, out=None, diag=None):
    """INTEGRATE_x for a batch of cells held in HOST numpy arrays.

    rconst [ncell,NREACT], fix [ncell,NFIX], var [ncell,NVAR] (not modified; with out=var the
    call advances var in place, as the C entry does).  `diag` = optional caller-owned
    (ierr, stats, hexit, texit) arrays, e.g. views of pinned memory so that the chunk pipeline
    of the library overlaps its copies.
    Returns (var_out, ierr[ncell], stats[ncell,8], hexit[ncell], texit[ncell])."""
    L = library(strict)
    nvar, nfix, nreact, _ = query(mech, strict)
    var_in = np.ascontiguousarray(var, dtype=np.float64).reshape(-1, nvar)
    ncell = var_in.shape[0]
    var_out = out if out is not None else np.empty_like(var_in)
    if var_out.shape != var_in.shape or var_out.dtype != np.float64 or not var_out.flags.c_contiguous:
        raise KppError("integrate: bad out array")
    if var_out.ctypes.data != var_in.ctypes.data:
        np.copyto(var_out, var_in)
    rconst = np.ascontiguousarray(rconst, dtype=np.float64).reshape(ncell, nreact)
    fix = np.ascontiguousarray(fix, dtype=np.float64).reshape(ncell, nfix)
    if diag is not None:
        ierr, stats, hexit, texit = diag
        for a, shp, dt in ((ierr, (ncell,), np.int32), (stats, (ncell, 8), np.int32),
                           (hexit, (ncell,), np.float64), (texit, (ncell,), np.float64)):
            if a.shape != shp or a.dtype != dt or not a.flags.c_contiguous:
                raise KppError("integrate: bad diag array (need contiguous %s %s)" % (np.dtype(dt).name, shp))
    else:
        ierr = np.zeros(ncell, dtype=np.int32)
        stats = np.zeros((ncell, 8), dtype=np.int32)
        hexit = np.zeros(ncell, dtype=np.float64)
        texit = np.zeros(ncell, dtype=np.float64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    o = opts if opts is not None else default_opts(strict)
    rc = L.mistra_kpp_integrate(mech, ncell, rconst.ctypes.data_as(dp), fix.ctypes.data_as(dp),
                                var_out.ctypes.data_as(dp), t0, t1, C.byref(o),
                                ierr.ctypes.data_as(ip), stats.ctypes.data_as(ip),
                                hexit.ctypes.data_as(dp), texit.ctypes.data_as(dp), None)
    _check(L, rc)
    return var_out, ierr, stats, hexit, texit


def integrate_device(mech, rconst, fix, var, t0=0.0, t1=10.0, opts=None, ierr=None, stats=None,
                     hexit=None, texit=None, stream=None, strict=False):
    """Same contract on DEVICE buffers: torch CUDA tensors (float64 / int32,
    contiguous) on the current device; `var` is advanced in place.  Asynchronous
    on `stream` (default: torch's current stream)."""
    import torch
    L = library(strict)
    nvar, nfix, nreact, _ = query(mech, strict)
    ncell = var.shape[0]
    for t, w, dt in ((rconst, nreact, torch.float64), (fix, nfix, torch.float64), (var, nvar, torch.float64)):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == dt and tuple(t.shape) == (ncell, w)):
            raise KppError("integrate_device: bad tensor (need contiguous CUDA %s [%d,%d])" % (dt, ncell, w))
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    o = opts if opts is not None else default_opts(strict)

    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None
    rc = L.mistra_kpp_integrate_device(mech, ncell, ptr(rconst), ptr(fix), ptr(var), t0, t1,
                                       C.byref(o), ptr(ierr), ptr(stats), ptr(hexit), ptr(texit),
                                       C.c_void_p(stream))
    _check(L, rc)


def integrate_multi(mech, rconst, fix, var, t0=0.0, t1=10.0, opts=None, devices=None, strict=False):
    """INTEGRATE_x for a batch of cells in HOST arrays over several GPUs of the box from this one process
    (mistra_kpp_integrate_multi): contiguous slices of cells, slice i on CUDA device devices[i] (None: all visible).
    Returns (var_out, ierr, stats, hexit, texit)."""
    L = library(strict)
    nvar, nfix, nreact, _ = query(mech, strict)
    var_out = np.array(var, dtype=np.float64, order="C").reshape(-1, nvar)
    ncell = var_out.shape[0]
    rconst = np.ascontiguousarray(rconst, dtype=np.float64).reshape(ncell, nreact)
    fix = np.ascontiguousarray(fix, dtype=np.float64).reshape(ncell, nfix)
    ierr = np.zeros(ncell, dtype=np.int32)
    stats = np.zeros((ncell, 8), dtype=np.int32)
    hexit = np.zeros(ncell, dtype=np.float64)
    texit = np.zeros(ncell, dtype=np.float64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    o = opts if opts is not None else default_opts(strict)
    if devices is None:
        ndev, dv = 0, None
    else:
        ndev = len(devices)
        dv = (C.c_int * ndev)(*devices)
    rc = L.mistra_kpp_integrate_multi(mech, ncell, rconst.ctypes.data_as(dp), fix.ctypes.data_as(dp),
                                      var_out.ctypes.data_as(dp), t0, t1, C.byref(o), ierr.ctypes.data_as(ip),
                                      stats.ctypes.data_as(ip), hexit.ctypes.data_as(dp), texit.ctypes.data_as(dp),
                                      ndev, dv)
    _check(L, rc)
    return var_out, ierr, stats, hexit, texit


def device_count(strict=False):
    return int(library(strict).mistra_kpp_device_count())


class PinnedArray:
    """A numpy array over page-locked host memory from mistra_kpp_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype=np.float64, strict=False):
        self._L = library(strict)
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        _check(self._L, self._L.mistra_kpp_host_alloc(C.byref(self._p), max(n, 1)))
        buf = (C.c_char * max(n, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._p:
                self.array = None
                self._L.mistra_kpp_host_free(self._p)
                self._p = None
        except Exception:
            pass


def host_register(a, strict=False):
    L = library(strict)
    _check(L, L.mistra_kpp_host_register(C.c_void_p(a.ctypes.data), a.nbytes))


def host_unregister(a, strict=False):
    L = library(strict)
    _check(L, L.mistra_kpp_host_unregister(C.c_void_p(a.ctypes.data)))


class RateList(C.Structure):
    _fields_ = [("n", C.c_int32), ("reserved", C.c_int32), ("idx", C.POINTER(C.c_int32)), ("val", C.POINTER(C.c_double))]


class RateInputsCompact(C.Structure):
    _fields_ = [("cb1", C.POINTER(C.c_double)), ("scal", C.POINTER(C.c_double)), ("ph_rat", C.POINTER(C.c_double)),
                ("ycw", C.POINTER(C.c_double)), ("ycwd", C.POINTER(C.c_double)),
                ("yhenry", RateList), ("yxkmt", RateList), ("ykef", RateList), ("ykeb", RateList),
                ("yxkmtd", RateList), ("yxeq", RateList), ("f32_literals", C.c_int32), ("reserved", C.c_int32)]


class CompactRates:
    """What Update_RCONST_x reads, in the compact form of include/mistra_kpp_rates.h: the per-layer scalars as they
    are, the KPP-species-indexed exchange arrays only for the species that are non-zero somewhere in the batch
    (one index list per array).  Built from the full arrays of include/mistra_rconst.h; keeps its buffers alive."""

    LISTS = ("yhenry", "yxkmt", "ykef", "ykeb", "yxkmtd", "yxeq")

    def __init__(self, mech, cb1, scal, ph_rat, ycw=None, ycwd=None, f32_literals=1, alloc=None, idx=None, **full):
        nvar, nfix, _, _ = query(mech)
        nspec = nvar + nfix
        self.mech = mech
        self.f32 = int(f32_literals)
        self.ncell = int(np.asarray(cb1).shape[0])
        alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype=dtype))

        def own(a, w):
            if a is None:
                return None
            out = alloc((self.ncell, w), np.float64)
            out[...] = np.asarray(a, dtype=np.float64).reshape(self.ncell, w)
            return out
        self.cb1, self.scal, self.ph_rat = own(cb1, 4), own(scal, 13), own(ph_rat, 47)
        nkc = 4 if mech == 2 else 2
        self.ycw, self.ycwd = own(ycw, nkc), own(ycwd, 2)
        self.idx, self.val = {}, {}
        for name in self.LISTS:
            a = full.get(name)
            if a is None:
                self.idx[name], self.val[name] = np.zeros(0, dtype=np.int32), None
                continue
            a = np.asarray(a, dtype=np.float64).reshape(self.ncell, -1, nspec)
            used = np.nonzero((a != 0.0).any(axis=(0, 1)))[0].astype(np.int32)
            if idx is not None:                      # a given list (e.g. of another slice of the same ensemble)
                if not set(used.tolist()) <= set(idx[name].tolist()):
                    raise KppError("CompactRates: %s has non-zero species outside the given index list" % name)
                used = np.ascontiguousarray(idx[name], dtype=np.int32)
            self.idx[name] = used
            v = alloc((self.ncell, a.shape[1], len(used)), np.float64)
            v[...] = a[:, :, used]
            self.val[name] = v
        self._bind()

    @classmethod
    def concat(cls, parts, alloc=None):
        """The batch made of the cells of `parts` (same mechanism, same index lists), buffers from `alloc`."""
        alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype=dtype))
        self = cls.__new__(cls)
        p0 = parts[0]
        self.mech, self.f32 = p0.mech, p0.f32
        self.ncell = sum(p.ncell for p in parts)

        def cat(arrs):
            if arrs[0] is None:
                return None
            out = alloc((self.ncell,) + arrs[0].shape[1:], np.float64)
            o = 0
            for a in arrs:
                out[o:o + a.shape[0]] = a
                o += a.shape[0]
            return out
        for k in ("cb1", "scal", "ph_rat", "ycw", "ycwd"):
            setattr(self, k, cat([getattr(p, k) for p in parts]))
        self.idx = {n: p0.idx[n].copy() for n in cls.LISTS}
        for p in parts[1:]:
            for n in cls.LISTS:
                if not np.array_equal(p.idx[n], p0.idx[n]):
                    raise KppError("CompactRates.concat: index lists differ")
        self.val = {n: cat([p.val[n] for p in parts]) for n in cls.LISTS}
        self._bind()
        return self

    def _bind(self):
        dp = C.POINTER(C.c_double)
        s = RateInputsCompact()
        for k in ("cb1", "scal", "ph_rat", "ycw", "ycwd"):
            a = getattr(self, k)
            setattr(s, k, a.ctypes.data_as(dp) if a is not None else None)
        for name in self.LISTS:
            rl = RateList()
            rl.n = len(self.idx[name])
            if rl.n:
                rl.idx = self.idx[name].ctypes.data_as(C.POINTER(C.c_int32))
                rl.val = self.val[name].ctypes.data_as(dp)
            setattr(s, name, rl)
        s.f32_literals = self.f32
        self.struct = s

    @property
    def nbytes(self):
        n = sum(a.nbytes for a in (self.cb1, self.scal, self.ph_rat, self.ycw, self.ycwd) if a is not None)
        return n + sum(v.nbytes for v in self.val.values() if v is not None) + sum(i.nbytes for i in self.idx.values())


def integrate_rates(mech, rates, fix, var, t0=0.0, t1=10.0, opts=None, out=None, diag=None):
    """mistra_kpp_integrate_rates: like integrate(), with the rate constants formed on the device from `rates`
    (a CompactRates).  Returns (var_out, ierr, stats, hexit, texit, h2d_bytes)."""
    L = library()
    nvar, nfix, nreact, _ = query(mech)
    var_in = np.ascontiguousarray(var, dtype=np.float64).reshape(-1, nvar)
    ncell = var_in.shape[0]
    if ncell != rates.ncell:
        raise KppError("integrate_rates: rates are for %d cells, var has %d" % (rates.ncell, ncell))
    var_out = out if out is not None else np.empty_like(var_in)
    if var_out.ctypes.data != var_in.ctypes.data:
        np.copyto(var_out, var_in)
    fix = np.ascontiguousarray(fix, dtype=np.float64).reshape(ncell, nfix)
    if diag is not None:
        ierr, stats, hexit, texit = diag
    else:
        ierr = np.zeros(ncell, dtype=np.int32)
        stats = np.zeros((ncell, 8), dtype=np.int32)
        hexit = np.zeros(ncell, dtype=np.float64)
        texit = np.zeros(ncell, dtype=np.float64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    o = opts if opts is not None else default_opts()
    moved = C.c_int64(0)
    L.mistra_kpp_integrate_rates.argtypes = [C.c_int, C.c_int64, C.POINTER(RateInputsCompact), dp, dp, C.c_double, C.c_double,
                                             C.POINTER(KppOpts), ip, ip, dp, dp, C.POINTER(C.c_int64), C.c_void_p]
    rc = L.mistra_kpp_integrate_rates(mech, ncell, C.byref(rates.struct), fix.ctypes.data_as(dp), var_out.ctypes.data_as(dp),
                                      t0, t1, C.byref(o), ierr.ctypes.data_as(ip), stats.ctypes.data_as(ip),
                                      hexit.ctypes.data_as(dp), texit.ctypes.data_as(dp), C.byref(moved), None)
    _check(L, rc)
    return var_out, ierr, stats, hexit, texit, int(moved.value)


def set_kernel(mech, variant, strict=False):
    """Kernel variant of a mechanism: -1 = by batch size (default), 0 = one cell per thread, 1 = on-chip (gas, aer)."""
    L = library(strict)
    _check(L, L.mistra_kpp_set_kernel(mech, variant))


def set_handoff(mech, steps, strict=False):
    """Step attempts a cell may make in the cell-per-thread kernel before the on-chip kernel continues it
    (include/mistra_kpp.h): -1 = default (12 for aer when the kernel is chosen by batch size), 0 = off."""
    L = library(strict)
    _check(L, L.mistra_kpp_set_handoff(mech, steps))


def handoff_count(strict=False):
    """Cells the last integrate_device call on the current device handed over to the on-chip kernel."""
    L = library(strict)
    L.mistra_kpp_handoff_count.restype = C.c_int64
    return int(L.mistra_kpp_handoff_count())


def get_kernel(mech, strict=False):
    """Pinned variant of a mechanism, -1 = chosen per call by the batch size."""
    return int(library(strict).mistra_kpp_get_kernel(mech))


def kernel_for(mech, ncell, strict=False):
    L = library(strict)
    L.mistra_kpp_kernel_for.argtypes = [C.c_int, C.c_int64]
    return int(L.mistra_kpp_kernel_for(mech, ncell))


def launch_count_variant(variant, strict=False):
    L = library(strict)
    L.mistra_kpp_launch_count_variant.restype = C.c_int64
    return int(L.mistra_kpp_launch_count_variant(variant))


def launch_count(strict=False):
    return int(library(strict).mistra_kpp_launch_count())


def fp64_peak_tflops():
    """Measured FP64 FMA peak of the current device (TFLOP/s)."""
    v = library().mistra_kpp_fp64_peak_tflops()
    if v <= 0:
        raise KppError("FP64 peak probe failed: no CUDA device?")
    return v


def finalize(strict=False):
    library(strict).mistra_kpp_finalize()
