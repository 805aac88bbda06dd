"""Binding of include/mistra_liq.h: the per-layer tables of the liq_parm chain (henry_x, v_mean_x, st_coeff_x,
equil_co_x; kpp.f90:664-2145, 2954-3363) on the host (libmistra_rconst.so) and on the device (libmistra_kpp.so)."""
import ctypes as C

import numpy as np

NSPEC = {1: 262, 2: 424}
NKC = {1: 2, 2: 4}
J6 = 55


class LiqArgs(C.Structure):
    _fields_ = [("t", C.c_void_p), ("conv2", C.c_void_p), ("xgamma", C.c_void_p), ("cw", C.c_void_p), ("cm", C.c_void_p),
                ("sion1_13_14", C.c_void_p), ("j6", C.c_int32), ("lpjoyce14bc", C.c_int32), ("f32_literals", C.c_int32),
                ("lpbuxmann15alph", C.c_int32), ("henry", C.c_void_p), ("vmean", C.c_void_p), ("alpha", C.c_void_p),
                ("xkef", C.c_void_p), ("xkeb", C.c_void_p)]


def tables_host(mech, t, conv2, xgamma, cw=None, cm=None, sion1_13_14=None, lpjoyce14bc=False, lpbuxmann15alph=False,
                f32_literals=1, nthreads=1):
    """Returns (henry, vmean, alpha [ncell][NSPEC], xkef, xkeb [ncell][nkc][NSPEC]) computed on the host."""
    from . import rconst
    L = rconst.lib()
    n, nkc, nspec = len(t), NKC[mech], NSPEC[mech]
    keep = []

    def arr(a, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == shape, (a.shape, shape)
        keep.append(a)
        return a.ctypes.data
    out = [np.empty((n, nspec)) for _ in range(3)] + [np.empty((n, nkc, nspec)) for _ in range(2)]
    a = LiqArgs(arr(t, (n,)), arr(conv2, (n, nkc)), arr(xgamma, (n, nkc, J6)), arr(cw, (n, nkc)), arr(cm, (n, nkc)),
                arr(sion1_13_14, (n, nkc, 2)), J6, int(lpjoyce14bc), int(f32_literals), int(lpbuxmann15alph),
                *[o.ctypes.data for o in out])
    L.mistra_liq_tables_host.argtypes = [C.c_int, C.c_int64, C.POINTER(LiqArgs), C.c_int]
    rc = L.mistra_liq_tables_host(mech, n, C.byref(a), nthreads)
    if rc != 0:
        raise ValueError("mistra_liq_tables_host failed (%d)" % rc)
    return tuple(out)


def tables_device(mech, t, conv2, xgamma, cw=None, cm=None, sion1_13_14=None, lpjoyce14bc=False, lpbuxmann15alph=False,
                  f32_literals=1, stream=None):
    """Same on the device: contiguous float64 CUDA tensors in, five CUDA tensors out (asynchronous on torch's stream)."""
    import torch
    from . import kpp
    L = kpp.library()
    n, nkc, nspec = t.shape[0], NKC[mech], NSPEC[mech]

    def ptr(x, shape):
        if x is None:
            return None
        if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float64 and tuple(x.shape) == shape):
            raise ValueError("liq.tables_device: need contiguous CUDA float64 %s" % (shape,))
        return x.data_ptr()
    out = [torch.empty((n, nspec), dtype=torch.float64, device=t.device) for _ in range(3)] + \
          [torch.empty((n, nkc, nspec), dtype=torch.float64, device=t.device) for _ in range(2)]
    a = LiqArgs(ptr(t, (n,)), ptr(conv2, (n, nkc)), ptr(xgamma, (n, nkc, J6)), ptr(cw, (n, nkc)), ptr(cm, (n, nkc)),
                ptr(sion1_13_14, (n, nkc, 2)), J6, int(lpjoyce14bc), int(f32_literals), int(lpbuxmann15alph),
                *[o.data_ptr() for o in out])
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    L.mistra_liq_tables_device.argtypes = [C.c_int, C.c_int64, C.POINTER(LiqArgs), C.c_void_p]
    kpp._check(L, L.mistra_liq_tables_device(mech, n, C.byref(a), C.c_void_p(stream)))
    return tuple(out)


def launch_count():
    from . import kpp
    L = kpp.library()
    L.mistra_liq_launch_count.restype = C.c_int64
    return int(L.mistra_liq_launch_count())
