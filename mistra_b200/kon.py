"""Host-side mirror of the reference's condensation step on the 2-D particle grid over the
C ABI of include/mistra_kon.h (libmistra_kpp.so).

`subkon` is SUBROUTINE subkon (/root/reference/src/str.f90:4987-5204, with advec
5321-5516) for a batch of humid layers - the call made per layer from SUBROUTINE kon
(str.f90:4705).  No CPU implementation lives here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import bins as _bins
from . import kpp as _kpp

MB, JPTAERRAD = 18, 3
R1 = 8.3144743 / 18.01528e-3          # constants.f90: r1
RHOW = 1000.0


class KonGrid(C.Structure):
    _fields_ = [("nka", C.c_int32), ("nkt", C.c_int32), ("a0m", C.c_double), ("dlne", C.c_double)] + [
        (n, C.POINTER(C.c_double)) for n in ("en", "rn", "b0m", "ew", "e", "dew", "rw", "qabs")] + [
        ("kw", C.POINTER(C.c_int32)), ("rq", C.POINTER(C.c_double)), ("ka", C.c_int32), ("reserved", C.c_int32)]


STATE_D = ("ff", "t", "talt", "xm1", "xm1a", "feu", "dfddt", "xm2", "dtcon", "p", "totrad")
SUMS = ("vol1_a", "vol1_d", "part_o_a", "part_o_d", "part_n_a", "part_n_d", "vol2", "pntot")


class KonState(C.Structure):
    """mistra_kon_state (include/mistra_kon.h)."""
    _fields_ = [(n, C.POINTER(C.c_double)) for n in STATE_D] + [("nar", C.POINTER(C.c_int32))] + [
        (n, C.POINTER(C.c_double)) for n in SUMS] + [("status", C.POINTER(C.c_int32))]


def kon_grid(rnw0=0.005, rnw1=15.0, rw0=0.005, rw1=150.0, nka=70, nkt=70, seed=7):
    """Particle grid of SUBROUTINE grid (via bins.particle_grid) plus what subkon reads from
    COMMON /cb44/ and /cb49/: the Koehler constants a0m = 152200/(r1*rhow) (str.f90:1317),
    b0m(ia) = fcs*xnue*xmol2/xmol3 (str.f90:1398; ammonium sulfate below ka, sea salt above)
    and an absorption-efficiency table qabs(18,nkt,nka,3).  The reference reads qabs from
    its Mie tables (input/*.dat, out of scope here); the synthetic one has the same shape
    and a physically plausible size/band dependence: Q_abs = 1 - exp(-k_band * r)."""
    g = _bins.particle_grid(rnw0, rnw1, rw0, rw1, nka, nkt)
    r = np.random.default_rng(seed)
    ka = g["ka"]
    fcs = np.where(np.arange(nka) < ka, r.uniform(0.5, 1.0, nka), 1.0)
    b0m = np.where(np.arange(nka) < ka, fcs * 3.0 * 18.0 / 132.0, fcs * 2.0 * 18.0 / 58.4)
    kband = np.concatenate([10.0 ** r.uniform(-4.0, -2.0, 6), 10.0 ** r.uniform(-1.5, -0.3, 12)])   # 1/um
    typ = np.array([1.0, 1.5, 0.7])
    qabs = 1.0 - np.exp(-typ[:, None, None, None] * kband[None, None, None, :] * g["rq"][None, :, :, None])
    g.update({"a0m": 152200.0 / (R1 * RHOW), "b0m": b0m, "qabs": np.ascontiguousarray(qabs)})
    return g


def _grid_struct(g):
    keep = [np.ascontiguousarray(g[n], dtype=np.float64) for n in ("en", "rn", "b0m", "ew", "e", "dew", "rw", "qabs")]
    assert keep[6].shape == (g["nka"], g["nkt"]) and keep[7].shape == (JPTAERRAD, g["nka"], g["nkt"], MB)
    kw = np.ascontiguousarray(g["kw"], dtype=np.int32)
    rq = np.ascontiguousarray(g["rq"], dtype=np.float64)
    s = KonGrid(g["nka"], g["nkt"], g["a0m"], g["dlne"], *[a.ctypes.data_as(C.POINTER(C.c_double)) for a in keep],
                kw.ctypes.data_as(C.POINTER(C.c_int32)), rq.ctypes.data_as(C.POINTER(C.c_double)), g["ka"], 0)
    return s, keep + [kw, rq]


def _lib(strict=False):
    L = _kpp.library(strict)
    if not getattr(L, "_kon_ready", False):
        L.mistra_kon_launch_count.restype = C.c_int64
        L._kon_ready = True
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def subkon(g, dt, ffk, totr, dfdt, feualt, pp, to, tn, xm1o, xm1n, kr):
    """HOST numpy arrays (not modified).  Returns (ffk, to, xm1o, status)."""
    L = _lib()
    gs, keep = _grid_struct(g)
    ffk = np.ascontiguousarray(ffk, dtype=np.float64).copy()
    n = ffk.shape[0]
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (totr, dfdt, feualt, pp, tn, xm1n)]
    to = np.ascontiguousarray(to, dtype=np.float64).copy()
    xm1o = np.ascontiguousarray(xm1o, dtype=np.float64).copy()
    kr = np.ascontiguousarray(kr, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    _kpp._check(L, L.mistra_kon_subkon(C.byref(gs), C.c_int64(n), C.c_double(dt), _dp(ffk), _dp(a[0]), _dp(a[1]),
                                       _dp(a[2]), _dp(a[3]), _dp(to), _dp(a[4]), _dp(xm1o), _dp(a[5]),
                                       kr.ctypes.data_as(C.POINTER(C.c_int32)),
                                       status.ctypes.data_as(C.POINTER(C.c_int32)), None))
    return ffk, to, xm1o, status


def subkon_device(g, dt, ffk, totr, dfdt, feualt, pp, to, tn, xm1o, xm1n, kr, status=None, stream=None):
    """DEVICE torch tensors (float64 / int32, contiguous); ffk, to, xm1o are updated in place;
    asynchronous on torch's current stream."""
    import torch
    L = _lib()
    gs, keep = _grid_struct(g)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream

    def vp(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None
    _kpp._check(L, L.mistra_kon_subkon_device(C.byref(gs), C.c_int64(ffk.shape[0]), C.c_double(dt), vp(ffk), vp(totr),
                                              vp(dfdt), vp(feualt), vp(pp), vp(to), vp(tn), vp(xm1o), vp(xm1n),
                                              vp(kr), vp(status), C.c_void_p(stream)))


def layers(g, dt, chem, st):
    """The layer loop of SUBROUTINE kon (str.f90:4615-4772) for a batch of layers held in HOST
    numpy arrays: `st` is a dict with ff, t, talt, xm1, xm1a, feu, dfddt, p, totrad, nar (not
    modified).  Returns a dict with the updated arrays plus xm2, dtcon, status and, with chem,
    vol1_a/d, part_o_a/d, part_n_a/d, vol2, pntot."""
    L = _lib()
    gs, keep = _grid_struct(g)
    n = st["ff"].shape[0]
    nka = g["nka"]
    o = {k: np.ascontiguousarray(st[k], dtype=np.float64).copy() for k in ("ff", "t", "talt", "xm1", "xm1a", "feu",
                                                                          "dfddt", "p", "totrad")}
    o["xm2"] = np.zeros(n); o["dtcon"] = np.zeros(n)
    o["nar"] = np.ascontiguousarray(st["nar"], dtype=np.int32)
    for k in SUMS[:6]:
        o[k] = np.zeros((n, nka))
    o["vol2"] = np.zeros((n, 4)); o["pntot"] = np.zeros((n, 4))
    o["status"] = np.zeros(n, dtype=np.int32)
    s = KonState(*[_dp(o[k]) for k in STATE_D], o["nar"].ctypes.data_as(C.POINTER(C.c_int32)),
                 *[(_dp(o[k]) if chem else None) for k in SUMS], o["status"].ctypes.data_as(C.POINTER(C.c_int32)))
    _kpp._check(L, L.mistra_kon_layers(C.byref(gs), C.c_int64(n), C.c_double(dt), C.c_int(1 if chem else 0),
                                       C.byref(s), None))
    return o


def synthetic_columns(g, ncell, seed=20261018, dry_fraction=0.3):
    """Inputs of the kon layer loop: the humid layers of synthetic_layers plus a share of
    dry layers (relative humidity 0.3..0.7, which take the Koehler-equilibrium branch)."""
    d = synthetic_layers(g, ncell, seed=seed)
    r = np.random.default_rng(seed + 1)
    dry = r.uniform(size=ncell) < dry_fraction
    rh = np.where(dry, r.uniform(0.3, 0.699, ncell), d["feualt"])
    es = p21(d["to"])
    xm1 = 0.62198 * rh * es / (d["pp"] - 0.37802 * rh * es)
    return {"ff": d["ffk"], "t": np.where(dry, d["to"], d["tn"]), "talt": d["to"],
            "xm1": np.where(dry, xm1, d["xm1n"]), "xm1a": np.where(dry, xm1, d["xm1o"]), "feu": rh,
            "dfddt": d["dfdt"], "p": d["pp"], "totrad": d["totr"], "nar": d["kr"]}


def launch_count():
    return int(_lib().mistra_kon_launch_count())


def p21(t):
    return 610.7 * np.exp(17.15 * (t - 273.15) / (t - 38.33))    # str.f90:7691


def synthetic_layers(g, ncell, seed=20261018):
    """Synthetic humid layers for subkon: temperature, pressure, relative humidity
    0.72..1.004, night or day radiative fluxes per band, and a bimodal particle spectrum
    whose classes sit at (or, when activated, beyond) their Koehler equilibrium size for the
    layer's humidity, spread over three water bins.  Returns a dict of numpy arrays."""
    r = np.random.default_rng(seed)
    nka, nkt = g["nka"], g["nkt"]
    rn, lrn = g["rn"], np.log(g["rn"])
    to = r.uniform(268.0, 295.0, ncell)
    pp = r.uniform(85.0e3, 101.0e3, ncell)
    rh = np.where(r.uniform(size=ncell) < 0.3, r.uniform(0.995, 1.004, ncell), r.uniform(0.72, 0.995, ncell))
    es = p21(to)
    xm1 = 0.62198 * rh * es / (pp - 0.37802 * rh * es)
    tn = to + r.uniform(-0.02, 0.02, ncell)
    xm1n = xm1 * (1.0 + r.uniform(-2.0e-4, 2.0e-4, ncell))
    dfdt = r.uniform(-2.0e-5, 2.0e-5, ncell)
    kr = r.integers(1, 4, ncell).astype(np.int32)
    day = r.uniform(size=ncell) < 0.5
    totr = np.concatenate([np.where(day[:, None], r.uniform(5.0, 60.0, (ncell, 6)), 0.0),
                           r.uniform(0.5, 5.0, (ncell, 12))], axis=1)
    ffk = np.zeros((ncell, nka, nkt))
    jdrop = int(np.searchsorted(g["rq"][nka // 2], 8.0))          # ~8 um droplets
    for c0 in range(0, ncell, 2048):
        m = min(2048, ncell - c0)
        sl = slice(c0, c0 + m)
        n1 = 10.0 ** r.uniform(2.0, 3.0, m); n2 = 10.0 ** r.uniform(-0.5, 0.8, m)
        mu1 = np.log(r.uniform(0.05, 0.12, m)); mu2 = np.log(r.uniform(0.8, 2.0, m))
        dist = (n1[:, None] * np.exp(-0.5 * ((lrn[None, :] - mu1[:, None]) / 0.55) ** 2)
                + n2[:, None] * np.exp(-0.5 * ((lrn[None, :] - mu2[:, None]) / 0.7) ** 2))
        dist *= (n1 + n2)[:, None] / dist.sum(axis=1, keepdims=True)
        a0 = g["a0m"] / to[sl]
        sr = np.exp(a0[:, None, None] / g["rw"][None] - (g["b0m"] * g["en"])[None, :, None] / g["ew"][None, None, :])
        above = sr >= rh[sl, None, None]
        jeq = np.where(above.any(axis=2), above.argmax(axis=2), min(nkt - 3, jdrop))
        jeq = np.clip(jeq, 1, nkt - 3)
        cc, ii = np.meshgrid(np.arange(m), np.arange(nka), indexing="ij")
        for dj, wgt in ((-1, 0.25), (0, 0.5), (1, 0.25)):
            ffk[c0 + cc, ii, jeq + dj] += wgt * dist
    return {"ffk": ffk, "totr": totr, "dfdt": dfdt, "feualt": rh, "pp": pp, "to": to, "tn": tn,
            "xm1o": xm1, "xm1n": xm1n, "kr": kr}
