"""ctypes binding of the CPU oracle (libkpp_oracle.so).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py - never from mistra_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double),
                ("hmin", C.c_double), ("hmax", C.c_double), ("hstart", C.c_double),
                ("facmin", C.c_double), ("facmax", C.c_double), ("facrej", C.c_double),
                ("facsafe", C.c_double),
                ("max_steps", C.c_int32), ("autonomous", C.c_int32),
                ("f32_literals", C.c_int32), ("reserved", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "libkpp_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.kpp_oracle_default_opts.argtypes = [C.POINTER(OracleOpts)]
        L.kpp_oracle_query.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4
        L.kpp_oracle_spc_name.argtypes = [C.c_int, C.c_int]
        L.kpp_oracle_spc_name.restype = C.c_char_p
        L.kpp_oracle_integrate.argtypes = [C.c_int, C.c_int64, dp, dp, dp, C.c_double, C.c_double,
                                           C.POINTER(OracleOpts), ip, ip, dp, dp, C.c_int]
        L.kpp_oracle_fun.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
        L.kpp_oracle_jac.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
        L.kpp_oracle_decomp.argtypes = [C.c_int, dp]
        L.kpp_oracle_solve.argtypes = [C.c_int, dp, dp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def default_opts(**kw):
    o = OracleOpts()
    lib().kpp_oracle_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def query(mech):
    v = [C.c_int() for _ in range(4)]
    assert lib().kpp_oracle_query(mech, *[C.byref(x) for x in v]) == 0
    return tuple(x.value for x in v)


def spc_name(mech, i):
    s = lib().kpp_oracle_spc_name(mech, i)
    return s.decode() if s else None


def integrate(mech, rconst, fix, var, t0=0.0, t1=10.0, opts=None, nthreads=1):
    """Returns (var_out, ierr, stats[ncell,8], hexit, texit); inputs untouched."""
    nvar, nfix, nreact, _ = query(mech)
    var = np.ascontiguousarray(var, dtype=np.float64).reshape(-1, nvar).copy()
    ncell = var.shape[0]
    rconst = np.ascontiguousarray(rconst, dtype=np.float64).reshape(ncell, nreact)
    fix = np.ascontiguousarray(fix, dtype=np.float64).reshape(ncell, nfix)
    ierr = np.zeros(ncell, dtype=np.int32)
    stats = np.zeros((ncell, 8), dtype=np.int32)
    hexit = np.zeros(ncell)
    texit = np.zeros(ncell)
    o = opts if opts is not None else default_opts()
    rc = lib().kpp_oracle_integrate(mech, ncell, _p(rconst), _p(fix), _p(var), t0, t1, C.byref(o),
                                    _ip(ierr), _ip(stats), _p(hexit), _p(texit), nthreads)
    assert rc == 0, rc
    return var, ierr, stats, hexit, texit


def fun(mech, V, F, RCT, f32=1):
    nvar = query(mech)[0]
    V, F, RCT = (np.ascontiguousarray(x, dtype=np.float64) for x in (V, F, RCT))
    out = np.zeros(nvar)
    lib().kpp_oracle_fun(mech, f32, _p(V), _p(F), _p(RCT), _p(out))
    return out


def jac(mech, V, F, RCT, f32=1):
    nz = query(mech)[3]
    V, F, RCT = (np.ascontiguousarray(x, dtype=np.float64) for x in (V, F, RCT))
    out = np.zeros(nz)
    lib().kpp_oracle_jac(mech, f32, _p(V), _p(F), _p(RCT), _p(out))
    return out


def decomp(mech, JVS):
    J = np.ascontiguousarray(JVS, dtype=np.float64).copy()
    ier = lib().kpp_oracle_decomp(mech, _p(J))
    return J, ier


def solve(mech, LU, X):
    LU = np.ascontiguousarray(LU, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64).copy()
    lib().kpp_oracle_solve(mech, _p(LU), _p(X))
    return X
