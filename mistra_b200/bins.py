"""Host-side mirror of the reference's 2-D bin redistribution around the chemistry
step, over the C ABI of include/mistra_bins.h (libmistra_kpp.so).

`snapshot` / `redistribute` are the two halves of SUBROUTINE stem_kpp
(/root/reference/src/str.f90:5916-5966 and 5976-6134) for a batch of layers;
`particle_grid` restates SUBROUTINE grid's 2-D particle grid (str.f90:1653-1705,
1882-1905).  No CPU implementation lives here: without the CUDA library or a device
every compute call raises."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import kpp as _kpp

NKC, LSP, J2, J6 = 4, 9, 121, 55
LJ2 = (1, 2, 8, 9, 13, 14, 19, 20, 30)          # str.f90:5876 (1-based ion indices)
ION_MASS = (1., 18., 96., 44., 62., 35.5, 97., 23., 95.)   # str.f90:5989-5993


class BinsGrid(C.Structure):
    _fields_ = [("nka", C.c_int32), ("nkt", C.c_int32), ("ka", C.c_int32), ("nkc_l", C.c_int32),
                ("ial_first", C.c_int32), ("reserved", C.c_int32),
                ("kw", C.POINTER(C.c_int32)), ("en", C.POINTER(C.c_double)), ("rq", C.POINTER(C.c_double))]


def particle_grid(rnw0=0.005, rnw1=15.0, rw0=0.005, rw1=150.0, nka=70, nkt=70, nkc_l=4,
                  chamber=False, ial_first=1):
    """The 2-D (dry aerosol mass x water mass) particle grid of SUBROUTINE grid
    (str.f90:1653-1705) and the bin limits ka, kw (str.f90:1882-1905) as a dict:
    en[nka] mg, e[nkt] mg, rn[nka] um, rq[nka,nkt] um, kw[nka], ka (1-based limits)."""
    pi = 3.1415926535897932
    rhow, rho3 = 1000.0, 2000.0
    x0 = 1.0 / 3.0
    x1 = 4.0 * x0 * pi * rhow
    x2 = 4.0 * x0 * pi * rho3
    enwmin = x2 * rnw0 ** 3 * 1.0e-12
    enwmax = x2 * rnw1 ** 3 * 1.0e-12
    x3 = 10.0 ** (math.log10(enwmax / enwmin) / nka)
    enw = np.empty(nka); en = np.empty(nka); rn = np.empty(nka)
    enw[0] = enwmin * x3
    en[0] = 0.5 * (enw[0] + enwmin)
    for ia in range(1, nka):
        enw[ia] = enw[ia - 1] * x3
        en[ia] = 0.5 * (enw[ia] + enw[ia - 1])
    rn[:] = (en / x2) ** x0 * 1.0e4
    ewmin = x1 * rw0 ** 3 * 1.0e-12
    ewmax = x1 * rw1 ** 3 * 1.0e-12
    ax = 10.0 ** (math.log10(ewmax / ewmin) / nkt)
    ew = np.empty(nkt); e = np.empty(nkt)
    ew[0] = ewmin * ax
    e[0] = 0.5 * (ew[0] + ewmin)
    for jt in range(1, nkt):
        ew[jt] = ew[jt - 1] * ax
        e[jt] = 0.5 * (ew[jt] + ew[jt - 1])
    rq = (e[None, :] * 1.0e-6 / x1 + (rn[:, None] * 1.0e-6) ** 3) ** x0 * 1.0e6       # [nka,nkt]
    rw = (ew[None, :] * 1.0e-6 / x1 + (rn[:, None] * 1.0e-6) ** 3) ** x0 * 1.0e6      # [nka,nkt]
    dew = np.empty(nkt)
    dew[0] = ew[0] - ewmin
    dew[1:] = ew[1:] - ew[:-1]
    dlgew = math.log10(ewmax / ewmin) / nkt
    zradthres = 0.1 if chamber else 0.5
    ka = -1
    for ia in range(nka):
        if rn[ia] > zradthres and ka < 0:
            ka = ia            # = (ia+1) - 1 in 1-based counting
    if ka < 0:
        ka = nka
    kw = np.full(nka, -1, dtype=np.int32)
    rwat = (e * 1.0e-6 / x1) ** x0 * 1.0e6
    for ia in range(nka):
        for jt in range(nkt):
            if rwat[jt] > 10.0 * rn[ia] and kw[ia] < 0:
                kw[ia] = jt    # = (jt+1) - 1
        if kw[ia] < 0:
            kw[ia] = nkt
    return {"nka": nka, "nkt": nkt, "ka": int(ka), "nkc_l": nkc_l, "ial_first": ial_first,
            "kw": kw, "en": en, "e": e, "rn": rn, "rq": np.ascontiguousarray(rq),
            "ew": ew, "dew": dew, "rw": np.ascontiguousarray(rw), "dlne": math.log(10.0) * dlgew}


def _grid_struct(g):
    keep = (np.ascontiguousarray(g["kw"], dtype=np.int32), np.ascontiguousarray(g["en"], dtype=np.float64),
            np.ascontiguousarray(g["rq"], dtype=np.float64))
    s = BinsGrid(g["nka"], g["nkt"], g["ka"], g["nkc_l"], g.get("ial_first", 1), 0,
                 keep[0].ctypes.data_as(C.POINTER(C.c_int32)), keep[1].ctypes.data_as(C.POINTER(C.c_double)),
                 keep[2].ctypes.data_as(C.POINTER(C.c_double)))
    return s, keep


def _lib(strict=False):
    L = _kpp.library(strict)
    if not getattr(L, "_bins_ready", False):
        L.mistra_bins_launch_count.restype = C.c_int64
        L._bins_ready = True
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def snapshot(g, ff, cm, sion1, sion1o=None, strict=False):
    """HOST numpy arrays.  Returns (sap[n,4], smp[n,4], sion1o[n,4,9]).  strict=True uses the
    -DKPP_STRICT build, which sums sap in the reference's running order (bit parity tests)."""
    L = _lib(strict)
    gs, keep = _grid_struct(g)
    ff = np.ascontiguousarray(ff, dtype=np.float64)
    n = ff.shape[0]
    assert ff.shape == (n, g["nka"], g["nkt"])
    cm = np.ascontiguousarray(cm, dtype=np.float64).reshape(n, NKC)
    sion1 = np.ascontiguousarray(sion1, dtype=np.float64).reshape(n, NKC, J6)
    sap = np.zeros((n, NKC)); smp = np.zeros((n, NKC))
    so = np.zeros((n, NKC, LSP)) if sion1o is None else np.ascontiguousarray(sion1o, dtype=np.float64).copy()
    _kpp._check(L, L.mistra_bins_snapshot(C.byref(gs), C.c_int64(n), _dp(ff), _dp(cm), _dp(sion1), _dp(sap),
                                          _dp(smp), _dp(so), None))
    return sap, smp, so


def redistribute(g, ff, cm, cw, sap, smp, sion1o, sion1, sl1, strict=False):
    """HOST numpy arrays (not modified).  Returns (ff, sion1, sl1, nwarn)."""
    L = _lib(strict)
    gs, keep = _grid_struct(g)
    ff = np.ascontiguousarray(ff, dtype=np.float64).copy()
    n = ff.shape[0]
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (cm, cw, sap, smp, sion1o)]
    sion1 = np.ascontiguousarray(sion1, dtype=np.float64).copy()
    sl1 = np.ascontiguousarray(sl1, dtype=np.float64).copy()
    nwarn = np.zeros(n, dtype=np.int32)
    _kpp._check(L, L.mistra_bins_redistribute(C.byref(gs), C.c_int64(n), _dp(ff), _dp(a[0]), _dp(a[1]), _dp(a[2]),
                                              _dp(a[3]), _dp(a[4]), _dp(sion1), _dp(sl1),
                                              nwarn.ctypes.data_as(C.POINTER(C.c_int32)), None))
    return ff, sion1, sl1, nwarn


def snapshot_device(g, ff, cm, sion1, sap, smp, sion1o, stream=None):
    """DEVICE torch tensors (float64, contiguous); asynchronous on torch's current stream."""
    import torch
    L = _lib()
    gs, keep = _grid_struct(g)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    _kpp._check(L, L.mistra_bins_snapshot_device(C.byref(gs), C.c_int64(ff.shape[0]), _vp(ff), _vp(cm), _vp(sion1),
                                                 _vp(sap), _vp(smp), _vp(sion1o), C.c_void_p(stream)))


def redistribute_device(g, ff, cm, cw, sap, smp, sion1o, sion1, sl1, nwarn=None, stream=None):
    import torch
    L = _lib()
    gs, keep = _grid_struct(g)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    _kpp._check(L, L.mistra_bins_redistribute_device(C.byref(gs), C.c_int64(ff.shape[0]), _vp(ff), _vp(cm), _vp(cw),
                                                     _vp(sap), _vp(smp), _vp(sion1o), _vp(sion1), _vp(sl1),
                                                     _vp(nwarn), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_bins_launch_count())


def synthetic_layers(g, ncell, seed=20261018, growth=0.05):
    """Synthetic inputs of the redistribution for `ncell` layers: a bimodal log-normal
    particle spectrum on the 2-D grid (sulfate + sea-salt mode, each particle at a
    humidity-dependent water mass), liquid water per chem bin from the spectrum (as cw_rc,
    kpp.f90:2152-2414, does), ion and dissolved-species loadings, and a post-chemistry
    ion state that differs by a few per cent (`growth`), so that every layer has both
    growing and shrinking bins.  Returns a dict of numpy arrays."""
    r = np.random.default_rng(seed)
    nka, nkt, ka = g["nka"], g["nkt"], g["ka"]
    rn = g["rn"]
    ff = np.zeros((ncell, nka, nkt))
    lrn = np.log(rn)
    for c0 in range(0, ncell, 4096):
        m = min(4096, ncell - c0)
        n1 = 10.0 ** r.uniform(2.0, 3.3, m)              # cm^-3, accumulation mode
        n2 = 10.0 ** r.uniform(-0.5, 1.0, m)             # cm^-3, coarse mode
        mu1 = np.log(r.uniform(0.05, 0.12, m)); mu2 = np.log(r.uniform(0.8, 2.0, m))
        s1, s2 = 0.55, 0.7
        dist = (n1[:, None] * np.exp(-0.5 * ((lrn[None, :] - mu1[:, None]) / s1) ** 2)
                + n2[:, None] * np.exp(-0.5 * ((lrn[None, :] - mu2[:, None]) / s2) ** 2))
        dist *= (n1 + n2)[:, None] / dist.sum(axis=1, keepdims=True)
        # water: every dry class sits around a wet/dry radius ratio (growth factor) 1.3..2.2,
        # spread over three neighbouring water bins; a fraction of the coarse mode is activated
        gf = r.uniform(1.3, 2.2, m)
        wet = rn[None, :] * gf[:, None]
        rq = g["rq"]                                      # [nka,nkt]
        jt0 = np.stack([np.searchsorted(rq[ia], wet[:, ia]) for ia in range(nka)], axis=1)
        jt0 = np.clip(jt0, 1, nkt - 2)
        cc, ii = np.meshgrid(np.arange(m), np.arange(nka), indexing="ij")
        for dj, wgt in ((-1, 0.25), (0, 0.5), (1, 0.25)):
            ff[c0 + cc, ii, jt0 + dj] += wgt * dist
        act = r.uniform(0.0, 0.3, m)                      # activated fraction of classes ia > ka
        kwv = np.asarray(g["kw"])
        for ia in range(ka, nka - 1):
            jd = min(nkt - 1, int(kwv[ia]) + 2)
            moved = ff[c0:c0 + m, ia, :] * act[:, None]
            ff[c0:c0 + m, ia, :] -= moved
            ff[c0:c0 + m, ia, jd] += moved.sum(axis=1)
    # liquid water content per chem bin [m3/m3] and the bin switches
    vol = 4.0 / 3.0 * np.pi * (g["rq"] * 1.0e-6) ** 3     # m3 per particle
    kwv = np.asarray(g["kw"])
    jt = np.arange(nkt)[None, :]
    aer_mask = jt < kwv[:, None]                          # [nka,nkt]
    small = (np.arange(nka) < ka)[:, None]
    masks = [small & aer_mask, (~small) & aer_mask, small & ~aer_mask, (~small) & ~aer_mask]
    cw = np.stack([(ff * (vol * mk)[None]).sum(axis=(1, 2)) * 1.0e6 for mk in masks], axis=1)   # per cm3 -> m3/m3
    cm = np.where(cw > 1.0e-13, cw * 1.0e3, 0.0)
    cm[:, g["nkc_l"]:] = 0.0
    # ions: the 9 mass-defining ions carry the bin's dry aerosol mass smp [mg cm^-3]
    # (sum_l sion1_l * M_l * 1e-6 * 1e3 = smp), the others are traces; dissolved gases sl1
    smp = np.stack([(ff * (g["en"][:, None] * mk)[None]).sum(axis=(1, 2)) for mk in masks], axis=1)
    sion1 = np.zeros((ncell, NKC, J6)); sl1 = np.zeros((ncell, NKC, J2))
    wgt = r.uniform(0.05, 1.0, (ncell, NKC, LSP))
    wgt /= wgt.sum(axis=2, keepdims=True)
    for l in range(J6):
        sion1[:, :, l] = r.uniform(0.0, 1.0, (ncell, NKC)) * 1.0e-4 * smp * 1.0e3 / 50.0
    for i, (l, mw) in enumerate(zip(LJ2, ION_MASS)):
        sion1[:, :, l - 1] = wgt[:, :, i] * smp * 1.0e3 / mw
    sl1[:] = r.uniform(0.0, 1.0, (ncell, NKC, J2)) * 1.0e-3 * (smp * 1.0e3 / 50.0)[:, :, None]
    sion1_new = sion1 * (1.0 + growth * r.uniform(-1.0, 1.0, (ncell, NKC, 1)) * r.uniform(0.5, 1.0, (ncell, NKC, J6)))
    return {"ff": ff, "cw": cw, "cm": cm, "sion1": sion1, "sion1_new": sion1_new, "sl1": sl1}
