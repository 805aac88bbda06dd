"""Host-side mirror of the gather / scatter halves of gas_drive / aer_drive / tot_drive
(/root/reference/src/gas.f:60-217, aer.f:59-246, tot.f:59-982) over the C ABI of include/mistra_drive.h:
copies between the model arrays s1, s3, sl1, sion1 and the KPP vectors (VAR, FIX) of a batch of layers,
on the device.  CUDA only - no CPU fallback."""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import kpp
from .mechgen import mech as mechmod

_HERE = os.path.dirname(os.path.abspath(__file__))
ARR = {"s1": 0, "s3": 1, "sl1": 2, "sion1": 3}


class DriveArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nvar", "nfix", "j1", "j5", "j2", "j6", "nkc", "nmap")] + [
        (n, C.c_void_p) for n in ("map_kpp", "map_arr", "map_off", "layer", "s1", "s3", "sl1", "sion1", "var", "fix")] + [
        ("clamp_liquid", C.c_int32), ("f32_literals", C.c_int32), ("indf_o2", C.c_int32), ("indf_h2o", C.c_int32),
        ("indf_n2", C.c_int32), ("indf_h2ol", C.c_int32 * 4)] + [(n, C.c_void_p) for n in ("air", "h2o", "cvv")] + [
        ("clip_negative", C.c_int32)]


def _lib():
    L = kpp.library()
    L.mistra_drive_gather_device.argtypes = [C.c_int64, C.POINTER(DriveArgs), C.c_void_p]
    L.mistra_drive_scatter_device.argtypes = [C.c_int64, C.POINTER(DriveArgs), C.c_void_p]
    L.mistra_drive_launch_count.restype = C.c_int64
    return L


def drive_map(mech_name, gas_names, rad_names, j2=121, j6=55, allow_unmatched=False):
    """The copy list of <mech>_drive: s1 / s3 entries matched by species name as match_mk_indexes does
    (utils.f90:84-140).  A user species name that the mechanism does not know aborts the reference
    (utils.f90:1052-1056); so it does here (ValueError) unless allow_unmatched=True, which leaves such names out - a
    misspelt name would otherwise be neither gathered nor scattered, silently.  sl1 / sion1 entries from
    aer_mk.dat / tot.f (mistra_b200/mech/drive_maps.json; none for the gas mechanism, whose six sl1
    reads - gas.f:151-156 - the caller appends).  Returns dict(kpp, arr, off) of int32 arrays and the FIX
    positions indf_*."""
    m = mechmod.load(mech_name)
    pos = {s: i + 1 for i, s in enumerate(m.spc_names)}
    missing = [s for s in list(gas_names) + list(rad_names) if s not in pos]
    if missing and not allow_unmatched:
        raise ValueError("drive_map: species not in mechanism %s: %s" % (mech_name, ", ".join(missing)))
    ent = [(pos[s], 0, j) for j, s in enumerate(gas_names) if s in pos]
    ent += [(pos[s], 1, j) for j, s in enumerate(rad_names) if s in pos]
    if mech_name in ("aer", "tot"):
        tab = json.load(open(os.path.join(_HERE, "mech", "drive_maps.json")))[mech_name]["gather"]
        for name, arr, j, kc in tab:
            row = j2 if arr == "sl1" else j6
            ent.append((pos[name], ARR[arr], (kc - 1) * row + j - 1))
    e = np.array(ent, dtype=np.int32).reshape(-1, 3)
    fixpos = lambda s: pos[s] - m.nvar if s in pos and pos[s] > m.nvar else 0
    return dict(kpp=np.ascontiguousarray(e[:, 0]), arr=np.ascontiguousarray(e[:, 1]), off=np.ascontiguousarray(e[:, 2]),
                indf_o2=fixpos("O2"), indf_h2o=fixpos("H2O"), indf_n2=fixpos("N2"),
                indf_h2ol=[fixpos("H2Ol%d" % b) for b in (1, 2, 3, 4)], nvar=m.nvar, nfix=m.nfix)


def _args(mp, layer, s1, s3, sl1, sion1, var, fix):
    import torch

    def ok(x, dt, nd):
        if not (x.is_cuda and x.is_contiguous() and x.dtype == dt and x.dim() == nd):
            raise ValueError("drive: need contiguous CUDA %s tensors of %d dimensions" % (dt, nd))
        return x.data_ptr()
    n = layer.numel()
    if tuple(var.shape) != (n, mp["nvar"]) or tuple(fix.shape) != (n, mp["nfix"]):
        raise ValueError("drive: var [ncell,NVAR], fix [ncell,NFIX]")
    nl, nkc = sl1.shape[0], sl1.shape[1]
    if sion1.shape[:2] != (nl, nkc) or s1.shape[0] != nl or s3.shape[0] != nl:
        raise ValueError("drive: s1 [nlayer,j1], s3 [nlayer,j5], sl1 [nlayer,nkc,j2], sion1 [nlayer,nkc,j6]")
    a = DriveArgs()
    a.nvar, a.nfix, a.j1, a.j5, a.j2, a.j6, a.nkc = mp["nvar"], mp["nfix"], s1.shape[1], s3.shape[1], sl1.shape[2], \
        sion1.shape[2], nkc
    a.nmap = int(mp["kpp_d"].numel())
    a.map_kpp, a.map_arr, a.map_off = (ok(mp[k], torch.int32, 1) for k in ("kpp_d", "arr_d", "off_d"))
    a.layer = ok(layer, torch.int64, 1)
    a.s1, a.s3, a.sl1, a.sion1 = ok(s1, torch.float64, 2), ok(s3, torch.float64, 2), ok(sl1, torch.float64, 3), \
        ok(sion1, torch.float64, 3)
    a.var, a.fix = ok(var, torch.float64, 2), ok(fix, torch.float64, 2)
    return a, n


def to_device(mp, device):
    """Uploads the map arrays of drive_map() once; returns the dict with kpp_d / arr_d / off_d tensors."""
    import torch
    out = dict(mp)
    for k in ("kpp", "arr", "off"):
        out[k + "_d"] = torch.from_numpy(mp[k]).to(device)
    return out


def gather_device(mp, layer, s1, s3, sl1, sion1, air, h2o, cvv, var, fix, clamp_liquid=True, f32_literals=True,
                  stream=None):
    """Fills var / fix [ncell, .] of the batch cells from the rows layer[i] of the model arrays (torch
    CUDA tensors; sl1 / sion1 are clamped at 0 in place as aer_drive does).  Asynchronous."""
    import torch
    L = _lib()
    a, n = _args(mp, layer, s1, s3, sl1, sion1, var, fix)
    a.clamp_liquid, a.f32_literals = int(clamp_liquid), int(f32_literals)
    a.indf_o2, a.indf_h2o, a.indf_n2 = mp["indf_o2"], mp["indf_h2o"], mp["indf_n2"]
    a.indf_h2ol = (C.c_int32 * 4)(*mp["indf_h2ol"])
    for name, x, shp in (("air", air, (n,)), ("h2o", h2o, (n,)), ("cvv", cvv, (n, 4))):
        if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float64 and tuple(x.shape) == shp):
            raise ValueError("drive: %s must be contiguous CUDA float64 %s" % (name, shp))
    a.air, a.h2o, a.cvv = air.data_ptr(), h2o.data_ptr(), cvv.data_ptr()
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_drive_gather_device(n, C.byref(a), C.c_void_p(stream)))


def scatter_device(mp, layer, s1, s3, sl1, sion1, var, fix, clip_negative=True, stream=None):
    """Copies the integrated concentrations back into the rows layer[i] of the model arrays and, with
    clip_negative, clips those rows at 0 as kpp_driver does (kpp.f90:4472-4477).  Asynchronous."""
    import torch
    L = _lib()
    a, n = _args(mp, layer, s1, s3, sl1, sion1, var, fix)
    a.clip_negative = int(clip_negative)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    kpp._check(L, L.mistra_drive_scatter_device(n, C.byref(a), C.c_void_p(stream)))


def launch_count():
    return int(_lib().mistra_drive_launch_count())
