#!/usr/bin/env python
"""Benchmark of the KPP chemistry hot path (BASELINE.json metric:
"KPP cell-integrations/sec (gas+aer) at 1/2/4/8 B200 vs host-CPU reference").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--cols C] [--mechs gas[,aer[,tot]]]

One "step" = one INTEGRATE call (t = 0 -> 10 s, reference options) over every cell
of the rank's synthetic ensemble (C columns x 148 cells of the gas
mechanism, plus the aqueous batches when requested), restarted from the same
saved spun-up state.  Weak scaling: every rank owns its own C columns; no
data-path collective, NCCL only reduces the diagnostics.

Prints ONE JSON line on rank 0 (contract in the task statement): `value` =
device-resident throughput, `e2e` = the same through the host-buffer C-ABI call
with H2D/D2H inside the timed region, `roofline` for the dominant kernel,
`cpu_baseline` = the CPU oracle (C restatement of the reference path; the
Fortran reference cannot be built in this image) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kpp_cell_integrations_per_s"
UNIT = "cell-integrations/s"
MECH_ID = {"gas": 0, "aer": 1, "tot": 2}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cols", type=int, default=int(os.environ.get("MISTRA_BENCH_COLS", "10000")),
                    help="columns per GPU (148 cells each)")
    ap.add_argument("--mechs", default=os.environ.get("MISTRA_BENCH_MECHS", "gas,aer"))
    ap.add_argument("--spinup", type=int, default=int(os.environ.get("MISTRA_BENCH_SPINUP", "12")))
    ap.add_argument("--cpu-sample-cols", type=int, default=200,
                    help="columns per step of the --impl reference arm")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --cols columns per GPU; strong: --cols columns in total, split over the GPUs")
    ap.add_argument("--tot-cells", type=int, default=int(os.environ.get("MISTRA_BENCH_TOT_CELLS", "100000")),
                    help="cloudy cells of the tot leg (SURVEY 8d: 1e5); 0 = skip")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the tot / cold-start / latency / on-chip / pageable legs (rank 0, N=1 only anyway)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bins", action="store_true", help="skip the 2-D particle-grid legs (bins, kon)")
    ap.add_argument("--kon-layers", type=int, default=int(os.environ.get("MISTRA_BENCH_KON_LAYERS", "10000")),
                    help="humid layers per GPU for the condensation leg (100 columns x 100)")
    ap.add_argument("--bins-layers", type=int, default=int(os.environ.get("MISTRA_BENCH_BINS_LAYERS", "29600")),
                    help="layers per GPU for the 2-D bin redistribution leg (200 columns x 148)")
    return ap.parse_args()


# ----------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------
def build_ensemble(mech, cols, rank, spinup, use_gpu, chunk_cols=500, want_rates=False):
    """Synthetic per-rank ensemble, spun up so that the timed step sees a stiff
    quasi-steady radical state.  Built in chunks of columns (the per-species input
    arrays of the aqueous mechanism are large).  Spin-up uses the CUDA path when a
    GPU is present (it is input preparation, not part of the timed region)."""
    from mistra_b200 import synthetic
    cls = {"gas": synthetic.GasEnsemble, "aer": synthetic.AerEnsemble, "tot": synthetic.TotEnsemble}.get(mech)
    if cls is None:
        raise SystemExit("bench: synthetic inputs for mechanism %r are not available yet" % mech)
    if use_gpu:
        from mistra_b200 import kpp

        def step(rc, fix, var):
            return kpp.integrate(MECH_ID[mech], rc, fix, var)[0]
    else:
        from oracle import kpp_oracle as ko

        def step(rc, fix, var):
            return ko.integrate(MECH_ID[mech], rc, fix, var, nthreads=os.cpu_count() or 1)[0]
    vs, rs, fs, crs = [], [], [], []
    for c0 in range(0, cols, chunk_cols):
        nc = min(chunk_cols, cols - c0)
        ens = cls(nc, col0=rank * cols + c0)
        var = ens.var
        rc = ens.rconst(var)
        for s in range(spinup):
            if s and s % 6 == 0:
                rc = ens.rconst(var)
            var = np.maximum(step(rc, ens.fix, var), 0.0)   # kpp_driver clips negatives (kpp.f90:4473-4477)
        if spinup > 0:
            rc = ens.rconst(var)
        vs.append(var); rs.append(rc); fs.append(ens.fix)
        if want_rates:
            crs.append(ens.compact_rates(idx=crs[0].idx if crs else None))
    return crs if want_rates else None, np.ascontiguousarray(np.concatenate(vs)), np.ascontiguousarray(np.concatenate(rs)), \
        np.ascontiguousarray(np.concatenate(fs))


def cpu_baseline(mech, var, rc, fix, ncells, threads):
    """CPU oracle (kind 'port': C restatement of the reference Fortran) on a
    bounded sample of the same workload."""
    from oracle import kpp_oracle as ko
    n = min(ncells, var.shape[0])
    t0 = time.perf_counter()
    _, ierr, stats, _, _ = ko.integrate(MECH_ID[mech], rc[:n], fix[:n], var[:n], nthreads=threads)
    dt = time.perf_counter() - t0
    return n / dt, dt, stats


def cpu_mixed(batches, frac, threads):
    """The same mechanism mix as the GPU step: the first `frac` of every batch.
    Returns (cells, seconds)."""
    cells, secs = 0, 0.0
    for mname, var, rc, fix in batches:
        n = max(1, int(var.shape[0] * frac))
        _, dt, _ = cpu_baseline(mname, var, rc, fix, n, threads)
        cells += n
        secs += dt
    return cells, secs


# ----------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, i.e.
    the oracle port (no Fortran compiler in this image), all host threads, on a
    bounded sample of the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    mechs = args.mechs.split(",")
    cols = args.cpu_sample_cols
    batches = []
    for mname in mechs:
        _, var, rc, fix = build_ensemble(mname, cols, 0, args.spinup, use_gpu=False, chunk_cols=50)
        batches.append((mname, var, rc, fix))
    n = sum(b[1].shape[0] for b in batches)
    for _ in range(args.warmup):
        cpu_mixed(batches, 1.0, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_mixed(batches, 1.0, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    mech = "+".join(mechs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "synthetic Mistra ensemble, %s mechanism, Ros3 0->10 s" % mech,
                   "sample": "%d columns = %d cells per step, same gas:aer mix as the b200 arm" % (cols, n)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d cells (%d columns), OpenMP over cells" % (n, cols)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ----------------------------------------------------------------------------
def parity_check(dbatches, batches, ncheck=1000):
    """A sample of the TIMED output (last timed step, device arm) against the CPU oracle on the same inputs:
    parity contract of tests/test_gpu_parity.py (RTOL 1e-3 with an ATOL floor of 1e-3 molecule cm^-3, identical
    ierr, share of cells on the oracle's accept/reject sequence)."""
    from oracle import kpp_oracle as ko
    out = {}
    ok_all = True
    for d, (mname, var, rc, fix) in zip(dbatches, batches):
        n = var.shape[0]
        idx = np.unique(np.linspace(0, n - 1, min(ncheck, n)).astype(np.int64))
        ref, ierr_o, stats_o, _, _ = ko.integrate(MECH_ID[mname], rc[idx], fix[idx], var[idx],
                                                  nthreads=os.cpu_count() or 1)
        tidx = torch_index(idx, d["var0"].device)
        got = d["vars"][-1][tidx].cpu().numpy()
        ierr = d["ierr"][tidx].cpu().numpy()
        stats = d["stats"][tidx].cpu().numpy()
        good = (ierr == 1) & (ierr_o == 1)
        rel = np.abs(got[good] - ref[good]) / (np.maximum(np.abs(got[good]), np.abs(ref[good])) + 1.66e-21)
        same = (stats[:, 2:5] == stats_o[:, 2:5]).all(axis=1)
        res = {"cells": int(len(idx)), "max_rel_err": float(rel.max()) if rel.size else 0.0,
               "ierr_equal": bool(np.array_equal(ierr, ierr_o)), "same_step_sequence": float(same.mean()),
               "ok": bool(np.array_equal(ierr, ierr_o) and (rel.size == 0 or rel.max() <= 1e-3))}
        ok_all = ok_all and res["ok"]
        out[mname] = res
    out["ok"] = ok_all
    out["note"] = ("last timed step of the device arm vs the CPU oracle (C restatement of the reference; parity "
                   "unpinned by the reference, SURVEY 8c) on evenly spaced cells; tolerance RTOL 1e-3, floor 1.66e-21 mol m^-3")
    return out


def torch_index(idx, device):
    import torch
    return torch.from_numpy(idx).to(device)


def time_device(kpp, torch, mech, d_rc, d_fix, d_var0, repeats=3):
    """Median device time (CUDA events) of one INTEGRATE over a device-resident batch, state restored outside the clock."""
    n = d_var0.shape[0]
    ierr = torch.zeros(n, dtype=torch.int32, device=d_var0.device)
    stats = torch.zeros((n, 8), dtype=torch.int32, device=d_var0.device)
    work = d_var0.clone()
    ts = []
    for it in range(repeats + 1):
        work.copy_(d_var0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kpp.integrate_device(mech, d_rc, d_fix, work, ierr=ierr, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        if it:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), stats.cpu().numpy(), ierr.cpu().numpy(), work


def run_extras(args, dev, batches, dbatches):
    """Legs beside the headline (rank 0, one GPU): the tot mechanism, a cold-start batch with rejected steps, the
    one-cell-per-call latency of the box-model configuration, the on-chip kernel variant, pageable host buffers."""
    import ctypes as C
    import torch
    from mistra_b200 import kpp, synthetic
    from mistra_b200.mechgen import mech as mechmod
    fp64_peak = kpp.fp64_peak_tflops()
    peaks, _ = measured_peaks()
    res = {}

    def roof(mname, stats, ms, ncell):
        m = mechmod.load(mname)
        flops = float(m.flops_from_stats(stats))
        nstp = float(stats[:, 2].sum())
        step_bytes = 8.0 * (8 * m.lu_nonzero + 40 * m.nvar + 4 * m.nreact)
        return {"cells": int(ncell), "kernel_ms": ms, "cells_per_s": ncell / (ms * 1e-3), "ros3_steps_per_s": nstp / (ms * 1e-3),
                "mean_steps_per_cell": nstp / max(1, ncell), "sum_nrej": float(stats[:, 4].sum()),
                "fp64": {"achieved": flops / (ms * 1e-3) * 1e-12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) * 1e-12 / fp64_peak},
                "hbm_stream": {"achieved": step_bytes * nstp / (ms * 1e-3) * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": step_bytes * nstp / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"]}}

    # ---- tot: 1e5 cloudy cells, spun up like the other batches ------------------------------------
    if args.tot_cells > 0:
        cols = max(1, args.tot_cells // synthetic.TotEnsemble.LAYERS)
        _, var, rc, fix = build_ensemble("tot", cols, 0, args.spinup, use_gpu=True, chunk_cols=250)
        ms, stats, ierr, _ = time_device(kpp, torch, 2, torch.from_numpy(rc).to(dev), torch.from_numpy(fix).to(dev),
                                         torch.from_numpy(var).to(dev))
        res["tot"] = roof("tot", stats, ms, var.shape[0])
        res["tot"]["failed_cells"] = int((ierr != 1).sum())
        res["tot"]["kernel"] = "ros3_kernel_t (cell per thread, workspace in HBM)"
        del var, rc, fix
        torch.cuda.empty_cache()

    # ---- cold start: no spin-up, radicals at zero -> many steps per cell and rejected steps -----------------------
    cold = {}
    for mname, cols in (("gas", 400), ("aer", 400)):
        _, var, rc, fix = build_ensemble(mname, cols, 0, 0, use_gpu=True)
        ms, stats, ierr, _ = time_device(kpp, torch, MECH_ID[mname], torch.from_numpy(rc).to(dev),
                                         torch.from_numpy(fix).to(dev), torch.from_numpy(var).to(dev), repeats=2)
        cold[mname] = roof(mname, stats, ms, var.shape[0])
        cold[mname]["failed_cells"] = int((ierr != 1).sum())
    cold["note"] = ("first chemistry step of a cold ensemble (radicals start at zero, kpp.f90:283-287): step counts differ "
                    "from cell to cell and steps are rejected, so lanes refill at different times")
    res["cold_start"] = cold

    # ---- on-chip kernel variant on the aer batch (first 98 000 cells) -------------------------------------------
    da = [d for d in dbatches if d["name"] == "aer"]
    if da:
        d = da[0]
        n = min(d["n"], 98000)
        kpp.set_kernel(1, 1)
        try:
            ms, stats, ierr, _ = time_device(kpp, torch, 1, d["rc"][:n].contiguous(), d["fix"][:n].contiguous(),
                                             d["var0"][:n].contiguous())
            kpp.set_kernel(1, 0)
            msd, statsd, _, _ = time_device(kpp, torch, 1, d["rc"][:n].contiguous(), d["fix"][:n].contiguous(),
                                            d["var0"][:n].contiguous())
        finally:
            kpp.set_kernel(1, -1)                      # back to the choice by batch size
        oc = roof("aer", stats, ms, n)
        oc["kernel"] = "ros3_onchip_a (one persistent block per SM, 5 cell slots, LU in shared memory / registers)"
        oc["cell_per_thread_kernel_same_cells"] = {"kernel_ms": msd, "cells_per_s": n / (msd * 1e-3)}
        oc["dram_bytes_per_cell"] = {"value": 14.7e3, "compulsory": 11984,
                                     "source": "ncu --set full capture profiles/r02_onchip_aer_ncu_full.txt (dram read+write / cells)"}
        res["onchip_aer"] = oc

    # ---- one cell per call: the box-model configuration namelist.Buys13_0D (kpp.f90:4296-4299) ----------------
    lat = {}
    host = os.path.join(ROOT, "tests", "host", "libb1_host.so")
    H = C.CDLL(host, mode=C.RTLD_GLOBAL) if os.path.exists(host) else None
    dp = C.POINTER(C.c_double)
    if H is not None:
        H.b1_latency_us.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_int]
        H.b1_latency_us.restype = C.c_double
    for mname, var, rc, fix in batches:
        v1, r1, f1 = (np.ascontiguousarray(a[:1]) for a in (var, rc, fix))
        mech = MECH_ID[mname]
        for _ in range(20):
            kpp.integrate(mech, r1, f1, v1)
        t0 = time.perf_counter()
        ncall = 300
        for _ in range(ncall):
            kpp.integrate(mech, r1, f1, v1)
        py_us = (time.perf_counter() - t0) / ncall * 1e6
        ent = {"b2_python_us_per_call": py_us}
        if H is not None:
            H.b1_latency_us(mech, v1.ctypes.data_as(dp), f1.ctypes.data_as(dp), r1.ctypes.data_as(dp), 50, 1)
            ent["b1_us_per_call"] = float(H.b1_latency_us(mech, v1.ctypes.data_as(dp), f1.ctypes.data_as(dp),
                                                          r1.ctypes.data_as(dp), 1000, 1))
        from oracle import kpp_oracle as ko
        t0 = time.perf_counter()
        for _ in range(50):
            ko.integrate(mech, r1, f1, v1, nthreads=1)
        ent["cpu_oracle_us_per_call"] = (time.perf_counter() - t0) / 50 * 1e6
        ent["kernel_variant"] = "on-chip" if kpp.kernel_for(mech, 1) == 1 else "cell per thread"
        lat[mname] = ent
    lat["note"] = ("one spun-up cell per call, 0 -> 10 s: B1 = integrate_x_ of libmistra_kpp_f77.so called from C with the "
                   "state in the COMMON blocks (tests/host/b1_host.c, 1000 calls), B2 = mistra_kpp_integrate through ctypes "
                   "(includes the Python call overhead), CPU = the oracle on one thread; 8640 such calls make up Buys13_0D")
    res["latency_1cell"] = lat
    return res


# ----------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from mistra_b200 import kpp
    from mistra_b200.mechgen import mech as mechmod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    mechs = args.mechs.split(",")
    cols_rank = args.cols
    if args.scaling == "strong":                       # the same --cols columns in total, whatever the GPU count
        cols_rank = args.cols // world + (1 if rank < args.cols % world else 0)
    batches = []
    rate_parts = {}
    for mname in mechs:
        crs, var, rc, fix = build_ensemble(mname, cols_rank, rank, args.spinup, use_gpu=True, want_rates=not args.no_e2e)
        rate_parts[mname] = crs
        batches.append((mname, var, rc, fix))
    ncell_rank = sum(b[1].shape[0] for b in batches)

    # ---- device-resident arm --------------------------------------------------
    stream = torch.cuda.current_stream()
    dbatches = []
    for mname, var, rc, fix in batches:
        n = var.shape[0]
        d = {
            "mech": MECH_ID[mname], "name": mname, "n": n,
            "var0": torch.from_numpy(var).to(dev), "var": torch.empty((n, var.shape[1]), dtype=torch.float64, device=dev),
            "rc": torch.from_numpy(rc).to(dev), "fix": torch.from_numpy(fix).to(dev),
            "ierr": torch.zeros(n, dtype=torch.int32, device=dev),
            "stats": torch.zeros((n, 8), dtype=torch.int32, device=dev),
            "hexit": torch.zeros(n, dtype=torch.float64, device=dev),
        }
        dbatches.append(d)

    kev = []  # (mech name, start event, end event) around each kernel launch of the timed steps

    # every timed step integrates its own copy of the saved state, made before the clock starts: the restore is
    # the bench's bookkeeping, not part of the path
    for d in dbatches:
        d["vars"] = [d["var0"].clone() for _ in range(args.steps)]

    def step_device(k):
        for d in dbatches:
            if k < 0:
                d["var"].copy_(d["var0"])
                v = d["var"]
            else:
                v = d["vars"][k]
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            kpp.integrate_device(d["mech"], d["rc"], d["fix"], v, 0.0, 10.0, ierr=d["ierr"],
                                 stats=d["stats"], hexit=d["hexit"])
            e1.record(stream)
            if k >= 0:
                kev.append((d["name"], e0, e1))

    for _ in range(args.warmup):
        step_device(-1)
    barrier()
    l0 = kpp.launch_count()
    with ClockSampler(local) as clk:
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record(stream)
        for k in range(args.steps):
            step_device(k)
        t_end.record(stream)
        barrier()
    launches = kpp.launch_count() - l0
    handed = torch.tensor([float(kpp.handoff_count())], dtype=torch.float64, device=dev)   # last call = the last mechanism
    if world > 1:
        dist.all_reduce(handed)
    ms_total = t_start.elapsed_time(t_end)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    tc = torch.tensor([float(ncell_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    total_cells = int(tc.item())
    value = total_cells / (ms_per_step * 1e-3)

    # diagnostics (the only collective on the path): sum of steps / rejects / failures
    diag = torch.zeros(4, dtype=torch.float64, device=dev)
    flops_rank = 0.0
    kernel_ms = {}
    for d in dbatches:
        st = d["stats"].double()
        diag[0] += st[:, 2].sum()
        diag[1] += st[:, 4].sum()
        diag[2] += (d["ierr"] != 1).sum()
        diag[3] += d["n"]
        m = mechmod.load(d["name"])
        d["flops"] = float(m.flops_from_stats(d["stats"].cpu().numpy()))
        flops_rank += d["flops"]
    for name, e0, e1 in kev:
        kernel_ms.setdefault(name, []).append(e0.elapsed_time(e1))
    if world > 1:
        dist.all_reduce(diag, op=dist.ReduceOp.SUM)

    # ---- roofline of the dominant kernel (rank 0's launches) -------------------
    dom = max(kernel_ms, key=lambda k: sum(kernel_ms[k]))
    dd = [d for d in dbatches if d["name"] == dom][0]
    kms = float(np.mean(kernel_ms[dom]))
    fp64_peak = kpp.fp64_peak_tflops()
    peaks, peak_src = measured_peaks()
    mdom = mechmod.load(dom)
    io_bytes = mdom.io_bytes * dd["n"]
    # HBM side: in this mapping (one cell per thread) the per-cell state streams through HBM,
    # so the kernel is HBM-bound.  Algorithmic bytes per Ros3 step:
    # (8 passes over the LU + 40 vector passes + 4 passes over RCONST) * 8 B
    step_bytes = 8.0 * (8 * mdom.lu_nonzero + 40 * mdom.nvar + 4 * mdom.nreact)
    nstp_dom = float(dd["stats"][:, 2].sum().item())
    # measured DRAM traffic per Ros3 step from the committed ncu --set full captures (profiles/)
    ncu_step_bytes = {"gas": 94.1e3, "aer": 565.6e3}.get(dom)   # profiles/r01_gas_*_ncu_full.txt, r01b_aer_ncu_full.txt
    roof = {
        "bound": "hbm", "kernel": "ros3_kernel_%s" % mdom.suffix,
        "achieved": step_bytes * nstp_dom / (kms * 1e-3) * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": step_bytes * nstp_dom / (kms * 1e-3) * 1e-9 / peaks["hbm_gbs"], "peak_source": peak_src,
        "kernel_ms": kms, "ros3_steps_per_launch": nstp_dom, "bytes_per_ros3_step": step_bytes,
        "traffic": (ncu_step_bytes * nstp_dom) if ncu_step_bytes else None,
        "traffic_source": "not measured in this run: dram read+write bytes per Ros3 step of the ncu --set full capture "
                          "profiles/r01b_aer_ncu_full.txt (aer) / profiles/r01_gas_cellperthread_ncu_full.txt (gas) x the Ros3 steps of this launch",
        "note": "algorithmic bytes = (8*LU_NONZERO + 40*NVAR + 4*NREACT)*8 B per Ros3 step x steps of the launch "
                "(workspace passes of the cell-per-thread mapping, SURVEY 8d); traffic = dram read+write per step "
                "of the ncu --set full capture in profiles/ x steps of this launch",
        "fp64": {"bound": "fp64", "achieved": dd["flops"] / (kms * 1e-3) * 1e-12, "peak": fp64_peak,
                 "unit": "TFLOP/s", "frac": dd["flops"] / (kms * 1e-3) * 1e-12 / fp64_peak,
                 "peak_source": "measured live: 8-chain DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 entry)",
                 "flops_per_launch": dd["flops"],
                 "note": "algorithmic flops of the reference formulation from the integrator's own counters "
                         "(SURVEY 8d; FMA=2, div=1; includes the dF/dT Fun call this kernel skips) - the north "
                         "star's target roofline, reachable only with the per-cell state on chip"},
        "compulsory_io": {"achieved": io_bytes / (kms * 1e-3) * 1e-9, "unit": "GB/s",
                          "note": "(2*NVAR+NFIX+NREACT)*8 B per cell-integration"},
    }

    # ---- end-to-end arm: host buffers through the C-ABI call --------------------
    e2e = None
    if not args.no_e2e:
        hb = []
        for mname, var, rc, fix in batches:
            hv0 = torch.from_numpy(var).pin_memory()
            nc = hv0.shape[0]
            dgp = (torch.empty(nc, dtype=torch.int32).pin_memory(), torch.empty((nc, 8), dtype=torch.int32).pin_memory(),
                   torch.empty(nc, dtype=torch.float64).pin_memory(), torch.empty(nc, dtype=torch.float64).pin_memory())
            hb.append((MECH_ID[mname], hv0, torch.empty_like(hv0).pin_memory(),
                       torch.from_numpy(rc).pin_memory(), torch.from_numpy(fix).pin_memory(), dgp))
        h2d = sum(v0.numel() * 8 + r.numel() * 8 + f.numel() * 8 for _, v0, _, r, f, _ in hb)
        d2h = sum(v0.numel() * 8 + v0.shape[0] * (4 + 32 + 8 + 8) for _, v0, _, r, f, _ in hb)

        def step_host(use_rates):
            # every buffer is pinned host memory; the saved state is restored outside the clock
            # (it is the bench's bookkeeping, not part of the path), the call itself moves
            # its inputs to the device and var, ierr, stats, hexit, texit back
            for mech, v0, v, r, f, dg in hb:
                v.copy_(v0)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            acc = 0
            moved = 0
            for (mech, v0, v, r, f, dg), cr in zip(hb, crp):
                vn = v.numpy()
                if use_rates:
                    _, ierr, stats, hexit, _tx, mv = kpp.integrate_rates(mech, cr, f.numpy(), vn, out=vn,
                                                                        diag=tuple(t.numpy() for t in dg))
                    moved += mv
                else:
                    _, ierr, stats, hexit, _tx = kpp.integrate(mech, r.numpy(), f.numpy(), vn, out=vn,
                                                               diag=tuple(t.numpy() for t in dg))
                acc += int((ierr != 1).sum())
            torch.cuda.synchronize()
            return time.perf_counter() - t0, moved

        def pinned(shape, dtype):
            return torch.empty(tuple(shape), dtype=torch.float64).pin_memory().numpy()
        crp = [kpp.CompactRates.concat(rate_parts[mname], alloc=pinned) for mname, _, _, _ in batches]
        rate_parts.clear()
        res = {}
        for use_rates in (False, True):
            for _ in range(max(1, args.warmup - 1)):
                step_host(use_rates)
            dt, moved = 0.0, 0
            for _ in range(args.steps):
                a, moved = step_host(use_rates)
                dt += a
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res[use_rates] = (float(tt.item()) / args.steps, moved)
        e2e = {"value": total_cells / res[True][0], "unit": UNIT,
               "h2d_bytes_per_step": int(res[True][1]), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": res[True][0] * 1e3, "host_buffers": "page-locked (pinned)",
               "path": "mistra_kpp_integrate_rates: compact inputs of Update_RCONST_x + FIX + VAR to the device, "
                       "expand -> Update_RCONST_x -> INTEGRATE_x there, VAR + diagnostics back",
               "rconst_path": {"value": total_cells / res[False][0], "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                               "ms_per_step": res[False][0] * 1e3,
                               "path": "mistra_kpp_integrate: RCONST evaluated by the host beforehand (outside the clock) and shipped"}}
        # the same call with PAGEABLE host arrays (what a Fortran caller has unless it registers its arrays with
        # mistra_kpp_host_register): copies are staged by the driver and do not overlap the kernels
        if world == 1 and not args.no_extras:
            pg = [(mech, v0.numpy().copy(), r.numpy().copy(), f.numpy().copy()) for mech, v0, v, r, f, dg in hb]
            dtp = []
            for it in range(3):
                work = [v0.copy() for _, v0, _, _ in pg]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for (mech, v0, r, f), w in zip(pg, work):
                    kpp.integrate(mech, r, f, w, out=w)
                torch.cuda.synchronize()
                dtp.append(time.perf_counter() - t0)
            e2e["pageable"] = {"value": total_cells / min(dtp[1:]), "unit": UNIT, "ms_per_step": min(dtp[1:]) * 1e3,
                               "note": "mistra_kpp_integrate (RCONST path) with plain numpy (pageable) arrays for rconst, fix, var and the diagnostics"}
            del pg, work

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        cpu_mixed(batches, 2048.0 / ncell_rank, threads)            # warm-up
        # bounded sample with the same mechanism mix: ~12 s of CPU work on all host threads
        # (cells are independent, so the rate scales linearly to the full batch)
        c_probe, t_probe = cpu_mixed(batches, min(1.0, 20000.0 / ncell_rank), threads)
        frac = min(1.0, 12.0 * (c_probe / t_probe) / ncell_rank)
        c_all, t_all = cpu_mixed(batches, frac, threads)
        frac1 = min(1.0, 4.0 * (c_probe / t_probe) / threads / ncell_rank)
        c_1, t_1 = cpu_mixed(batches, frac1, 1)
        cpu = {"value": c_all / t_all, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %.3g of every batch (%d cells, mechanisms %s), OpenMP over cells, %.1f s"
                         % (frac, c_all, "+".join(mechs), t_all),
               "value_1thread": c_1 / t_1, "sample_1thread": "%d cells, %.1f s" % (c_1, t_1),
               "note": "C restatement of the reference Fortran (no Fortran compiler in the image)"}

    parity = parity_check(dbatches, batches) if rank == 0 else None
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        for d in dbatches:
            d["vars"] = d["vars"][-1:]                 # free the per-step copies
        torch.cuda.empty_cache()
        extras = run_extras(args, dev, batches, dbatches)
    bins_res = None
    kon_res = None
    next_res = None
    if not args.no_bins:
        bins_res = run_bins_leg(args, dev, world, rank, barrier)
        kon_res = run_kon_leg(args, dev, world, rank, barrier)
        next_res = run_next_rows_leg(args, dev, world, rank, barrier)

    if rank == 0:
        clocks = clk.summary()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic ensemble of independent Mistra columns (SURVEY 8d), "
                                   "%d columns %s (148 gas cells, 98 aer / tot cells per column), mechanisms: %s; Ros3 0->10 s, "
                                   "RTOL 1e-3 ATOL 1e-25 Hstart 1e-3" % (args.cols, "per GPU" if args.scaling == "weak" else "in total", "+".join(mechs)),
                       "cells_per_gpu": ncell_rank, "columns_per_gpu": cols_rank,
                       "spinup": "%d chemistry steps of 10 s before the timed state is saved (SURVEY 8d asks for 60; the mean "
                                 "Ros3 steps per cell and the share of rejected steps no longer change after ~10, DESIGN.md 7)" % args.spinup,
                       "timed_region": "K INTEGRATE calls per mechanism, each on its own copy of the saved state (copies made before the clock starts)",
                       "parallelism": "cells sharded by column over %d GPU(s), no data-path collective" % world,
                       "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2; no explicit flush"
                             % (sum(d["rc"].numel() * 8 + d["var"].numel() * 8 for d in dbatches) * 1e-9)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "handoff": {"cells_continued_on_chip_last_call": int(handed.item()),
                        "note": "mistra_kpp_set_handoff default: an aer cell that has made 12 step attempts in the cell-per-thread "
                                "kernel is continued by the on-chip kernel (one extra launch per call, counted in gpu_launches); "
                                "value / roofline time the whole call, hand-off pass included; MISTRA_KPP_HANDOFF=0 switches it off"},
            "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "tot": extras.get("tot") if extras else None,
            "cold_start": extras.get("cold_start") if extras else None,
            "onchip_aer": extras.get("onchip_aer") if extras else None,
            "latency_1cell": extras.get("latency_1cell") if extras else None,
            "diagnostics": {"sum_nstp": float(diag[0]), "sum_nrej": float(diag[1]),
                            "failed_cells": float(diag[2]), "cells": float(diag[3]),
                            "mean_steps_per_cell": float(diag[0] / max(1.0, float(diag[3])))},
            "bins": bins_res, "kon": kon_res, "next_rows": next_res,
            "per_mechanism": {k: {"kernel_ms": float(np.mean(v)),
                                  "cells_per_s": [d["n"] for d in dbatches if d["name"] == k][0] / (np.mean(v) * 1e-3)}
                              for k, v in kernel_ms.items()},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_bins_leg(args, dev, world, rank, barrier):
    """Second hot-path row: the 2-D bin redistribution around the chemistry call
    (stem_kpp, str.f90:5916-6134).  One step = snapshot + redistribute over every layer.
    Reported beside the headline, not part of `value`."""
    import torch
    import torch.distributed as dist
    from mistra_b200 import bins
    grid = bins.particle_grid()
    n = args.bins_layers
    d = bins.synthetic_layers(grid, n, seed=20261018 + rank)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in d.items()}
    ff0, si0, sl0 = t["ff"].clone(), t["sion1_new"].clone(), t["sl1"].clone()
    sap = torch.zeros((n, 4), dtype=torch.float64, device=dev)
    smp = torch.zeros_like(sap)
    so = torch.zeros((n, 4, 9), dtype=torch.float64, device=dev)
    nw = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    ev = []

    def step(timed):
        t["ff"].copy_(ff0); t["sion1_new"].copy_(si0); t["sl1"].copy_(sl0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        bins.snapshot_device(grid, t["ff"], t["cm"], t["sion1"], sap, smp, so)
        e1.record(stream)
        bins.redistribute_device(grid, t["ff"], t["cm"], t["cw"], sap, smp, so, t["sion1_new"], t["sl1"], nw)
        e2.record(stream)
        if timed:
            ev.append((e0, e1, e2))
    for _ in range(args.warmup):
        step(False)
    barrier()
    l0 = bins.launch_count()
    for _ in range(args.steps):
        step(True)
    barrier()
    launches = bins.launch_count() - l0
    ms_snap = float(np.mean([a.elapsed_time(b) for a, b, _ in ev]))
    ms_red = float(np.mean([b.elapsed_time(c) for _, b, c in ev]))
    tt = torch.tensor([ms_snap + ms_red], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    peaks, peak_src = measured_peaks()
    tile = grid["nka"] * grid["nkt"] * 8
    bytes_red = n * (2 * tile + 2 * 4 * (121 + 55) * 8)       # ff + sl1 + sion1, read and written once
    bytes_snap = n * tile
    res = {"metric": "bin_redistribution_layers_per_s", "value": n * world / (ms * 1e-3), "unit": "layers/s",
           "layers_per_gpu": n, "ms_per_step": ms, "gpu_launches": int(launches),
           "workload": "synthetic 70x70 particle spectra, 4 chem bins, +-5 %% ion mass change per bin "
                       "(%.1f GB of ff per GPU: larger than L2)" % (n * tile * 1e-9),
           "roofline": {"bound": "hbm", "kernel": "bins_redistribute_kernel",
                        "achieved": bytes_red / (ms_red * 1e-3) * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": bytes_red / (ms_red * 1e-3) * 1e-9 / peaks["hbm_gbs"], "peak_source": peak_src,
                        "kernel_ms": ms_red, "traffic": None,
                        "note": "algorithmic bytes per layer = 2*(nka*nkt + 4*(j2+j6))*8 = %d B" % (bytes_red // n),
                        "snapshot": {"kernel": "bins_snapshot_kernel", "kernel_ms": ms_snap,
                                     "achieved": bytes_snap / (ms_snap * 1e-3) * 1e-9,
                                     "frac": bytes_snap / (ms_snap * 1e-3) * 1e-9 / peaks["hbm_gbs"]}}}
    if not args.no_e2e:
        h = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in d.items()}
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            s_, m_, o_ = bins.snapshot(grid, h["ff"], h["cm"], h["sion1"])
            bins.redistribute(grid, h["ff"], h["cm"], h["cw"], s_, m_, o_, h["sion1_new"], h["sl1"])
        dt = (time.perf_counter() - t0) / reps
        res["e2e"] = {"value": n * world / dt, "unit": "layers/s", "ms_per_step": dt * 1e3,
                      "h2d_bytes_per_step": int(2 * n * tile + n * 8 * (4 * 4 + 2 * 4 * 55 + 2 * 4 * 9 + 4 * 121)),
                      "d2h_bytes_per_step": int(n * tile + n * 8 * (2 * 4 + 4 * 9 + 4 * 55 + 4 * 121) + 4 * n)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import bins_oracle as bo
        m, reps, t0 = n, 0, time.perf_counter()
        while time.perf_counter() - t0 < 5.0:                  # ~5 s of CPU work
            s_, m_, o_ = bo.snapshot(grid, d["ff"][:m], d["cm"][:m], d["sion1"][:m])
            bo.redistribute(grid, d["ff"][:m], d["cm"][:m], d["cw"][:m], s_, m_, o_, d["sion1_new"][:m], d["sl1"][:m])
            reps += 1
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": m * reps / dt, "unit": "layers/s", "cores": 1, "kind": "port",
                               "sample": "%d x %d layers, single thread (as the reference runs), %.1f s" % (reps, m, dt)}
    return res


def run_kon_leg(args, dev, world, rank, barrier):
    """Third hot-path row: condensation / evaporation on the 2-D particle grid (subkon +
    advec, str.f90:4987-5204, 5321-5516), one step = one subkon call over every layer.
    Reported beside the headline, not part of `value`."""
    import torch
    import torch.distributed as dist
    from mistra_b200 import kon
    from mistra_b200 import kpp
    grid = kon.kon_grid()
    n = args.kon_layers
    d = kon.synthetic_layers(grid, n, seed=20261018 + rank)
    keys = ("ffk", "totr", "dfdt", "feualt", "pp", "to", "tn", "xm1o", "xm1n", "kr")
    t0 = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev) for k in keys}
    t = {k: v.clone() for k, v in t0.items()}
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    ev = []

    def step(timed):
        for k in ("ffk", "to", "xm1o"):
            t[k].copy_(t0[k])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kon.subkon_device(grid, 10.0, *[t[k] for k in keys], status=st)
        e1.record(stream)
        if timed:
            ev.append((e0, e1))
    for _ in range(args.warmup):
        step(False)
    barrier()
    l0 = kon.launch_count()
    for _ in range(args.steps):
        step(True)
    barrier()
    launches = kon.launch_count() - l0
    tt = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ev]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    iters = float(st.clamp(min=0).float().mean().item())
    nk = grid["nka"] * grid["nkt"]
    # algorithmic flops per layer (FMA = 2): coefficient set-up ~ (60 + 5*18*2) per grid point incl.
    # the band sum, per secant iteration ~ 110 per grid point (two c(jt), u, 4th-order fluxes, dwsum)
    flops = n * nk * (60.0 + 5.0 * 18.0 * 2.0 + iters * 110.0)
    peak = kpp.fp64_peak_tflops()
    peaks, peak_src = measured_peaks()
    res = {"metric": "condensation_layers_per_s", "value": n * world / (ms * 1e-3), "unit": "layers/s",
           "layers_per_gpu": n, "ms_per_step": ms, "gpu_launches": int(launches), "mean_iterations": iters,
           "workload": "synthetic humid layers (RH 0.72..1.004), 70x70 particle spectra near Koehler "
                       "equilibrium, dt = 10 s (%.1f GB of ff per GPU)" % (n * nk * 8e-9),
           "roofline": {"bound": "fp64", "kernel": "kon_subkon_kernel", "achieved": flops / (ms * 1e-3) * 1e-12,
                        "peak": peak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) * 1e-12 / peak,
                        "kernel_ms": ms, "traffic": None,
                        "note": "one CTA (512 threads) per SM, five 70x70 tiles in shared memory; the FP64 pipe "
                                "mostly executes IEEE divisions (30 per grid point in the set-up, ~10 per flux), "
                                "which the flop count above books as 1 flop each; "
                                "HBM side: 2*nka*nkt*8 = %d B per layer -> %.4f of the measured copy bandwidth"
                                % (2 * nk * 8, n * 2 * nk * 8 / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"])}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import kon_oracle as ko
        m, reps, t1 = n, 0, time.perf_counter()
        while time.perf_counter() - t1 < 5.0:                  # ~5 s of CPU work
            ko.subkon(grid, 10.0, *[d[k][:m] for k in keys])
            reps += 1
        dt = time.perf_counter() - t1
        res["cpu_baseline"] = {"value": m * reps / dt, "unit": "layers/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d x %d layers, OpenMP over layers, %.1f s" % (reps, m, dt)}
    return res


def run_next_rows_leg(args, dev, world, rank, barrier):
    """Rows N1 and N3 of SURVEY 8f, each timed on its own: Update_RCONST_a on the device
    (gas.f:275-666 / aer.f:304-1400) and konc (kpp.f90:3370-3585).  Reported beside the headline."""
    import torch
    from mistra_b200 import konc, kpp, rconst as rcm, synthetic
    peaks, peak_src = measured_peaks()
    stream = torch.cuda.current_stream()
    res = {}

    def timeit(fn, restore=None):
        ev = []
        for i in range(args.warmup + args.steps):
            if restore:
                restore()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            if i >= args.warmup:
                ev.append((e0, e1))
        barrier()
        return float(np.mean([a.elapsed_time(b) for a, b in ev]))

    # ---- konc ----
    n = args.bins_layers * 8                                     # 0.24 M layers: 3.5 GB of sums + species
    d = konc.synthetic_sums(n, seed=20261018 + rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sums = {k: t(v) for k, v in d["sums"].items()}
    vol2, pntot, sl0, si0 = t(d["vol2"]), t(d["pntot"]), t(d["sl1"]), t(d["sion1"])
    sl1, sion1 = sl0.clone(), si0.clone()
    l0 = konc.launch_count()
    ms = timeit(lambda: konc.konc_device(d["ka"], sums, vol2, pntot, sl1, sion1),
                restore=lambda: (sl1.copy_(sl0), sion1.copy_(si0)))
    by = n * (2 * 4 * (konc.J2 + konc.J6) + 6 * 70 + 8) * 8
    res["konc"] = {"metric": "konc_layers_per_s", "value": n * world / (ms * 1e-3), "unit": "layers/s",
                   "layers_per_gpu": n, "ms_per_step": ms, "gpu_launches": int(konc.launch_count() - l0),
                   "roofline": {"bound": "hbm", "kernel": "konc_kernel", "achieved": by / (ms * 1e-3) * 1e-9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                "peak_source": peak_src, "traffic": None,
                                "note": "algorithmic bytes per layer = (2*4*(j2+j6) + 6*nka + 8)*8 = %d B" % (by // n)}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import konc_oracle as kco
        m = min(n, 20000)
        sub = {k: v[:m] for k, v in d["sums"].items()}
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            kco.konc(d["ka"], sub, d["vol2"][:m], d["pntot"][:m], d["sl1"][:m], d["sion1"][:m])
            reps += 1
        dt = time.perf_counter() - t1
        res["konc"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "layers/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "%d x %d layers, OpenMP over layers, %.1f s" % (reps, m, dt)}
    del sums, vol2, pntot, sl0, si0, sl1, sion1
    # ---- cw_rc ----
    from mistra_b200 import cwrc, kon
    g = kon.kon_grid()
    n = args.bins_layers
    st = kon.synthetic_columns(g, n, seed=20261018 + rank, dry_fraction=0.3)
    ffh = st["ff"] * 100.0
    cloudh = np.ones((n, 4), dtype=np.int32)
    gd = {"nka": g["nka"], "nkt": g["nkt"], "ka": g["ka"], "kw": t(np.asarray(g["kw"], dtype=np.int32)),
          "e": t(g["e"]), "rq": t(g["rq"])}
    ffd, feud, cloudd = t(ffh), t(st["feu"]), t(cloudh)
    outs = [torch.empty((n, 4), dtype=torch.float64, device=dev) for _ in range(4)]
    l0 = cwrc.launch_count()
    ms = timeit(lambda: cwrc.cw_rc_device(gd, ffd, feud, cloudd, *outs))
    by = n * (g["nka"] * g["nkt"] * 8 + 8 + 16 + 4 * 32)
    res["cw_rc"] = {"metric": "cw_rc_layers_per_s", "value": n * world / (ms * 1e-3), "unit": "layers/s",
                    "layers_per_gpu": n, "ms_per_step": ms, "gpu_launches": int(cwrc.launch_count() - l0),
                    "roofline": {"bound": "hbm", "kernel": "cwrc_kernel", "achieved": by / (ms * 1e-3) * 1e-9,
                                 "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                 "peak_source": peak_src, "traffic": None,
                                 "note": "algorithmic bytes per layer = nka*nkt*8 + 152 = %d B (%.1f GB of ff: larger than L2)"
                                         % (by // n, n * g["nka"] * g["nkt"] * 8e-9)}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cwrc_oracle as cwo
        m = min(n, 8000)
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            cwo.cw_rc(g, ffh[:m], st["feu"][:m], cloudh[:m])
            reps += 1
        dt = time.perf_counter() - t1
        res["cw_rc"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "layers/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "%d x %d layers, OpenMP over layers, %.1f s" % (reps, m, dt)}
    # ---- fast_k_mt_a (same ff; cw and cm as cw_rc just wrote them) ----
    from mistra_b200 import fastkmt
    rr = np.random.default_rng(20261018 + rank)
    ns, nx = 262, 50
    lexd = t(fastkmt.lex("aer"))
    tt, pp = t(st["t"]), t(st["p"])
    freepd = t(2.28e-5 * st["t"] / st["p"])
    alphad = t(10.0 ** rr.uniform(-4, 0, (n, ns)))
    vmeand = t(rr.uniform(100.0, 700.0, (n, ns)))
    xkd = torch.zeros((n, 4, ns), dtype=torch.float64, device=dev)
    vtd = torch.zeros((n, 4), dtype=torch.float64, device=dev)
    fk = {}
    for tag, ffx in (("synthetic", ffd), ("dense", None)):
        if ffx is None:                                          # every grid point populated: the FP64-bound case
            ffx = ffd + 1.0e-3
            cwrc.cw_rc_device(gd, ffx, feud, cloudd, *outs)
        l0 = fastkmt.launch_count()
        ms = timeit(lambda: fastkmt.fast_k_mt_device(gd, lexd, ffx, freepd, tt, pp, outs[1], outs[2], alphad, vmeand,
                                                      xkd, vtd))
        torch.cuda.synchronize()
        on = (outs[2] > 0).cpu().numpy()                         # bins with chemistry
        kcg = np.where(np.arange(g["nkt"])[None, :] < np.asarray(g["kw"])[:, None], 0, 2) + \
            np.where(np.arange(g["nka"])[:, None] < g["ka"], 0, 1)
        pop = (ffx > 0).cpu().numpy()
        npt = sum(int((pop[on[:, kc]] & (kcg == kc)[None]).sum()) for kc in range(4))   # points that enter the sums
        flops = npt * nx * 7.0 + n * g["nka"] * g["nkt"] * 1.0   # per term: add, div, 4 mul, add (div = 1 flop); q = rqm/freep
        by = n * (g["nka"] * g["nkt"] * 8 + 2 * ns * 8 + 11 * 8) + int(on.sum()) * nx * 8
        fk[tag] = {"value": n * world / (ms * 1e-3), "ms_per_step": ms, "gpu_launches": int(fastkmt.launch_count() - l0),
                   "populated_points_per_layer": npt / n, "bins_with_chemistry_per_layer": float(on.sum()) / n,
                   "fp64": {"achieved": flops / (ms * 1e-3) * 1e-12, "unit": "TFLOP/s",
                            "divisions_per_s": npt * nx / (ms * 1e-3)},
                   "hbm": {"achieved": by / (ms * 1e-3) * 1e-9, "unit": "GB/s",
                           "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"]}}
    p64 = kpp.fp64_peak_tflops()
    for v in fk.values():
        v["fp64"]["peak"] = p64
        v["fp64"]["frac"] = v["fp64"]["achieved"] / p64
    res["fast_k_mt"] = {"metric": "fast_k_mt_layers_per_s", "value": fk["synthetic"]["value"], "unit": "layers/s",
                        "layers_per_gpu": n, "ms_per_step": fk["synthetic"]["ms_per_step"],
                        "gpu_launches": fk["synthetic"]["gpu_launches"], "mechanism": "aer",
                        "roofline": {"bound": "hbm", "kernel": "fastkmt_kernel", "achieved": fk["synthetic"]["hbm"]["achieved"],
                                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": fk["synthetic"]["hbm"]["frac"],
                                     "peak_source": peak_src, "traffic": None,
                                     "note": "synthetic spectra populate ~4 % of the grid, so the layer is read faster than "
                                             "it is integrated: HBM-side bytes per layer = nka*nkt*8 + 2*NSPEC*8 + 88 + 400 per "
                                             "bin with chemistry; 'dense' = every grid point populated, FP64-bound: (nx + 1) "
                                             "IEEE divisions per point, booked as 1 flop each"},
                        "cases": fk}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import fastkmt_oracle as fko
        m = min(n, 2000)
        h = lambda x: x[:m].cpu().numpy()
        cwrc.cw_rc_device(gd, ffd, feud, cloudd, *outs)
        torch.cuda.synchronize()
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            fko.fast_k_mt(g, fastkmt.lex("aer"), ffh[:m], h(freepd), h(tt), h(pp), h(outs[1]), h(outs[2]), h(alphad),
                          h(vmeand), h(xkd), h(vtd))
            reps += 1
        dt = time.perf_counter() - t1
        res["fast_k_mt"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "layers/s", "cores": os.cpu_count(),
                                            "kind": "port", "sample": "%d x %d layers (synthetic spectra; the reference "
                                            "loop visits every grid point), OpenMP over layers, %.1f s" % (reps, m, dt)}
    del ffd, feud, cloudd, outs, alphad, vmeand, xkd, vtd
    # ---- difc (row N4, first piece): implicit vertical exchange of every chemical species ----
    from mistra_b200 import difc as dm
    ncol, nlev = max(1, args.cols // 5), 150
    dc = dm.synthetic_columns(ncol, nlev, seed=20261018 + rank)
    rows = ((93, 93), (24, 24), (121 * 4, 121 * 4), (55 * 4, 55 * 4))    # s1, s3, sl1, sion1 of the reference
    rr = np.random.default_rng(20261018 + rank)
    fd = [(torch.from_numpy(rr.uniform(0.5, 1.5, (ncol, nlev, r))).to(dev) * t(dc["am3"])[:, :, None], p) for r, p in rows]
    dd = {k: t(dc[k]) for k in ("atkh", "w", "am3", "detw", "deta")}
    l0 = dm.launch_count()
    ms = timeit(lambda: dm.difc_device(60.0, dd["atkh"], dd["w"], dd["am3"], dd["detw"], dd["deta"], fd))
    nsp = sum(p for _, p in rows)
    by = ncol * nsp * (2 * (nlev - 2) + 1) * 8 + ncol * nlev * 3 * 8
    res["difc"] = {"metric": "difc_columns_per_s", "value": ncol * world / (ms * 1e-3), "unit": "columns/s",
                   "columns_per_gpu": ncol, "levels": nlev, "species_per_column": nsp, "ms_per_step": ms,
                   "gpu_launches": int(dm.launch_count() - l0),
                   "tridiagonal_systems_per_s": ncol * nsp * world / (ms * 1e-3),
                   "roofline": {"bound": "hbm", "kernel": "difc_solve_kernel", "achieved": by / (ms * 1e-3) * 1e-9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                "peak_source": peak_src, "traffic": None,
                                "note": "algorithmic bytes per species and column = (2*(n-2) + 1)*8 (every level read "
                                        "once and written once) + 3*n*8 per column; %.1f GB of species arrays: larger "
                                        "than L2; the four solve launches and the coefficient launch are timed together"
                                        % (ncol * nsp * nlev * 8e-9)}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import difc_oracle as dfo
        m = min(ncol, 200)
        fh = [(a[:m].cpu().numpy(), p) for a, p in fd]
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            dfo.difc(60.0, dc["atkh"][:m], dc["w"][:m], dc["am3"][:m], dc["detw"], dc["deta"], fh)
            reps += 1
        dt = time.perf_counter() - t1
        res["difc"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "columns/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "%d x %d columns, OpenMP over columns (includes a copy of the arrays), "
                                                 "%.1f s" % (reps, m, dt)}
    del fd
    # ---- difp: the same exchange for the 70 x 70 particle spectrum of every level ----
    ncp = max(1, ncol // 4)                                        # 500 columns: 2.9 GB of ff
    rho = t(1.2 * np.exp(-np.cumsum(dc["detw"])[None] / 8000.0) * np.ones((ncp, 1)))
    ffp = torch.rand((ncp, nlev, 4900), dtype=torch.float64, device=dev)
    fsp = torch.zeros((ncp, nlev), dtype=torch.float64, device=dev)
    l0 = dm.launch_count()
    ms = timeit(lambda: dm.difp_device(60.0, dd["atkh"][:ncp], dd["w"][:ncp], rho, dd["detw"], dd["deta"], ffp, fsp))
    by = ncp * 4900 * 2 * (nlev - 1) * 8
    res["difp"] = {"metric": "difp_columns_per_s", "value": ncp * world / (ms * 1e-3), "unit": "columns/s",
                   "columns_per_gpu": ncp, "levels": nlev, "ms_per_step": ms, "gpu_launches": int(dm.launch_count() - l0),
                   "tridiagonal_systems_per_s": ncp * 4900 * world / (ms * 1e-3),
                   "roofline": {"bound": "hbm", "kernel": "difc_solve_kernel<true>", "achieved": by / (ms * 1e-3) * 1e-9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                "peak_source": peak_src, "traffic": None,
                                "note": "algorithmic bytes per grid point and column = 2*(n-1)*8: every level read and "
                                        "written once (the level sums fsum ride along in the backward sweep); %.1f GB of ff: "
                                        "larger than L2; coefficient, solve and fsum launches timed together"
                                        % (ncp * nlev * 4900 * 8e-9)}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = min(ncp, 32)
        fh, sh, rh = ffp[:m].cpu().numpy(), fsp[:m].cpu().numpy(), rho[:m].cpu().numpy()
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            dfo.difp(60.0, dc["atkh"][:m], dc["w"][:m], rh, dc["detw"], dc["deta"], fh, sh)
            reps += 1
        dt = time.perf_counter() - t1
        res["difp"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "columns/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "%d x %d columns, OpenMP over columns (includes a copy of the arrays), "
                                                 "%.1f s" % (reps, m, dt)}
    del dd, ffp, fsp, rho
    # ---- sedp / sedl: gravitational settling of the spectrum and of the aqueous species (row N4, the rest) ----
    from mistra_b200 import kon as konm, sed as sedm
    from oracle import sed_oracle as sdo
    ncs, nfl = max(1, ncol // 8), 100                              # 250 columns: 1.5 GB of ff
    sg = konm.kon_grid()
    sd = sedm.synthetic_columns(sg, ncs, n=nlev, nf=nfl, seed=20261018 + rank)
    sgd = dict(nka=sg["nka"], nkt=sg["nkt"], rq=t(sg["rq"]), e=t(sg["e"]), kw=t(np.asarray(sg["kw"], dtype=np.int32)))
    sv = {k: t(v) for k, v in sd.items()}
    sk = {k: sv[k].clone() for k in ("ff", "diag", "sl1", "sion1")}
    settle = float(((sd["ff"][:, 1:nfl] * sd["detw"][None, 1:nfl, None, None]).sum(axis=1) > 1e-6).mean())
    l0 = sedm.launch_count()
    ms = timeit(lambda: sedm.sedp_device(sgd, 10.0, nfl, sv["detw"], sv["deta"], sv["t"], sv["p"], sv["vd"], sv["ff"], sv["diag"]),
                restore=lambda: (sv["ff"].copy_(sk["ff"]), sv["diag"].copy_(sk["diag"])))
    by = ncs * (nfl - 1) * 4900 * 8 * (1.0 + 2.0 * settle)
    res["sedp"] = {"metric": "sedp_columns_per_s", "value": ncs * world / (ms * 1e-3), "unit": "columns/s",
                   "columns_per_gpu": ncs, "levels": nfl, "settling_classes": settle, "ms_per_step": ms,
                   "gpu_launches": int(sedm.launch_count() - l0),
                   "roofline": {"bound": "hbm", "kernel": "sedp_scan_kernel + sedp_work_kernel", "achieved": by / (ms * 1e-3) * 1e-9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                "peak_source": peak_src, "traffic": None,
                                "note": "algorithmic bytes = every class and level 2..nf read once for the column sums, the "
                                        "settling classes (%.1f %%) read once more and written once; their ~270 FP64 "
                                        "instructions per level and sub-step (six IEEE divisions) make the work kernel "
                                        "latency / FP64-bound" % (100 * settle)}}
    l0 = sedm.launch_count()
    ms = timeit(lambda: sedm.sedl_device(10.0, nfl, 4, sv["detw"], sv["deta"], sv["t"], sv["p"], sv["rc"], sv["vt"], sv["vdm"],
                                         sv["sl1"], sv["sion1"]),
                restore=lambda: (sv["sl1"].copy_(sk["sl1"]), sv["sion1"].copy_(sk["sion1"])))
    by = ncs * (nfl - 1) * 4 * (121 + 55) * 8 * 2
    res["sedl"] = {"metric": "sedl_columns_per_s", "value": ncs * world / (ms * 1e-3), "unit": "columns/s",
                   "columns_per_gpu": ncs, "levels": nfl, "profiles_per_column": 4 * (121 + 55), "ms_per_step": ms,
                   "gpu_launches": int(sedm.launch_count() - l0),
                   "roofline": {"bound": "hbm", "kernel": "sedl_kernel", "achieved": by / (ms * 1e-3) * 1e-9,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                "peak_source": peak_src, "traffic": None,
                                "note": "algorithmic bytes = levels 2..nf of sl1 and sion1 read and written once; every profile "
                                        "runs advsed1 (six IEEE divisions per level and sub-step): FP64 / latency-bound"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = min(ncs, 16)
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            sdo.sedp(sg, 10.0, nfl, sd["detw"], sd["deta"], sd["t"][:m], sd["p"][:m], sd["vd"][:m], sd["ff"][:m], sd["diag"][:m])
            reps += 1
        dt = time.perf_counter() - t1
        res["sedp"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "columns/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "%d x %d columns, OpenMP over columns (includes a copy of ff), %.1f s" % (reps, m, dt)}
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            for arr in (sd["sl1"], sd["sion1"]):
                sdo.sedl(10.0, nfl, 4, sd["detw"], sd["deta"], sd["t"][:m], sd["p"][:m], sd["rc"][:m], sd["vt"][:m], sd["vdm"][:m], arr[:m])
            reps += 1
        dt = time.perf_counter() - t1
        res["sedl"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "columns/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "%d x %d columns, OpenMP over columns, %.1f s" % (reps, m, dt)}
    del sv, sk, sd
    # ---- the layer loop of kpp_driver (row a15): per-layer scalars, switches and the layers sorted by mechanism ----
    from mistra_b200 import driver as drv
    from oracle import driver_oracle as dvo
    ncd_, nfd = max(1, args.cols // 5), 100                         # 2000 columns x 150 levels = 300 000 layers
    dcol = drv.synthetic_columns(ncd_, nlev, nfd, 4, 93, 24, 20261018 + rank)
    dcol["u0"] = np.random.default_rng(20261018 + rank).uniform(-0.2, 1.0, ncd_)
    cfg = dict(nf=nfd, halo=True, iod=True, lpBuys13_0D=False, neula=0, box=False, n_bl=2, kinv=70, dt_ch=10.0)
    adv_row, xadv = np.array([4, -1, 17, 60], dtype=np.int32), np.array([1e-9, 2e-9, -3e-10, 1e-8])
    dvd = {k: t(v) for k, v in dcol.items() if k != "cloud"}
    dvd["cloud"] = torch.from_numpy(dcol["cloud"].astype(np.int32)).to(dev)
    nl = ncd_ * nlev
    zz = lambda *sh, dt=torch.float64: torch.zeros(sh, dtype=dt, device=dev)
    dout = dict(cb1=zz(nl, 4), scal=zz(nl, 13), ph_rat=zz(nl, 47), air=zz(nl), h2o=zz(nl), cvv=zz(nl, 4),
                mech=zz(nl, dt=torch.int32), layers=zz(3, nl, dt=torch.int64), count=zz(3, dt=torch.int64))
    advr, xadvd = torch.from_numpy(adv_row).to(dev), t(xadv)
    l0 = drv.launch_count()
    ms = timeit(lambda: drv.layers_device(cfg, dvd["u0"], dvd["t"], dvd["p"], dvd["rho"], dvd["cm3"], dvd["am3"], dvd["xm1"],
                                          dvd["conv2"], dvd["cm"], dvd["cloud"], dvd["photol_j"], dout, advr, xadvd,
                                          dvd["s1"], dvd["s3"]))
    by = nl * ((6 + 2 * 4 + 2 * 47) * 8 + 4 * 4 + 2 * (93 + 24) * 8 + (4 + 13 + 47 + 2 + 4) * 8 + 4 + 4 + 8)
    res["driver"] = {"metric": "kpp_driver_layers_per_s", "value": nl * world / (ms * 1e-3), "unit": "layers/s",
                     "layers_per_gpu": nl, "ms_per_step": ms, "gpu_launches": int(drv.launch_count() - l0),
                     "layers_by_mechanism": [int(x) for x in dout["count"].cpu().numpy()],
                     "roofline": {"bound": "hbm", "kernel": "driver_layer_kernel + clip / list kernels", "achieved": by / (ms * 1e-3) * 1e-9,
                                  "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                  "peak_source": peak_src, "traffic": None,
                                  "note": "algorithmic bytes per layer = inputs (6 scalars, conv2, cm, cloud, two levels of photol_j) + s1 / s3 "
                                          "read and written by the clip + the output rows + mech read twice by the list kernels = %d B" % (by // nl)}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = min(ncd_, 200)
        cut = {k: (v[:m] if k not in ("detw", "deta") else v) for k, v in dcol.items()}
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 2.0:
            dvo.layers(cfg, cut["u0"], cut["t"], cut["p"], cut["rho"], cut["cm3"], cut["am3"], cut["xm1"], cut["conv2"], cut["cm"],
                       cut["cloud"], cut["photol_j"], adv_row, xadv, cut["s1"], cut["s3"])
            reps += 1
        dt = time.perf_counter() - t1
        res["driver"]["cpu_baseline"] = {"value": m * nlev * reps / dt, "unit": "layers/s", "cores": 1, "kind": "port",
                                         "sample": "%d x %d layers, numpy restatement (vectorised over layers), %.1f s" % (reps, m * nlev, dt)}
    del dvd, dout, dcol
    # ---- gather / scatter halves of aer_drive on device-resident model arrays (rows a14 / a15) ----
    from mistra_b200 import drive
    from mistra_b200.mechgen import mech as mechmod
    ma = mechmod.load("aer")
    gasph = [x for x in ma.spc_names[:ma.nvar] if x[-2:] not in ("l1", "l2")]
    mp = drive.to_device(drive.drive_map("aer", gasph[:75], gasph[75:95]), dev)
    ncd = max(1, args.cols // 10) * 98                            # the aer cells of the rconst leg
    rr = np.random.default_rng(20261018 + rank)
    lay = torch.from_numpy(rr.permutation(ncd).astype(np.int64)).to(dev)
    marr = [torch.rand(sh, dtype=torch.float64, device=dev) for sh in ((ncd, 75), (ncd, 20), (ncd, 4, 121), (ncd, 4, 55))]
    vard = torch.zeros((ncd, ma.nvar), dtype=torch.float64, device=dev)
    fixd = torch.zeros((ncd, ma.nfix), dtype=torch.float64, device=dev)
    aird, h2od, cvvd = (torch.rand(sh, dtype=torch.float64, device=dev) + 0.5 for sh in ((ncd,), (ncd,), (ncd, 4)))
    l0 = drive.launch_count()

    def drive_step():
        drive.gather_device(mp, lay, *marr, aird, h2od, cvvd, vard, fixd)
        drive.scatter_device(mp, lay, *marr, vard, fixd)
    ms = timeit(drive_step)
    nm = int(mp["kpp_d"].numel())
    by = ncd * ((2 * nm + 7) * 8 * 2 + 4 * (121 + 55) * 16 * 2 + (75 + 20) * 16)
    res["drive"] = {"metric": "drive_gather_scatter_cells_per_s", "value": ncd * world / (ms * 1e-3), "unit": "cells/s",
                    "mechanism": "aer", "cells_per_gpu": ncd, "ms_per_step": ms, "gpu_launches": int(drive.launch_count() - l0),
                    "roofline": {"bound": "hbm", "kernel": "drive_gather_kernel + drive_scatter_kernel",
                                 "achieved": by / (ms * 1e-3) * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"], "peak_source": peak_src, "traffic": None,
                                 "note": "per cell: %d copies each way (read + write), the clamp of the liquid rows (4*176 "
                                         "doubles read + written) before and the clip of all four rows after; layers in "
                                         "random order" % nm}}
    del marr, vard, fixd, lay
    # ---- Update_RCONST_a ----
    ens = synthetic.AerEnsemble(max(1, args.cols // 10), seed=20261018 + rank)
    tn = lambda a: None if a is None else t(a)
    fields = ("yhenry", "yxkmt", "ykef", "ykeb", "yxkmtd", "yxeq", "ycw", "ycwd")
    kw = {k: tn(getattr(ens, k, None)) for k in fields}
    cb1, scal, ph, conc = t(ens.cb1), t(ens.scal), t(ens.ph_rat), t(ens.conc())
    out = torch.empty((ens.ncell, 979), dtype=torch.float64, device=dev)
    ms = timeit(lambda: rcm.update_rconst_device(1, cb1, scal, ph, conc, out=out, **kw))
    inb = sum(v.numel() for v in [cb1, scal, ph, conc] + [x for x in kw.values() if x is not None]) * 8
    by = inb + out.numel() * 8
    res["rconst"] = {"metric": "update_rconst_cells_per_s", "value": ens.ncell * world / (ms * 1e-3), "unit": "cells/s",
                     "mechanism": "aer", "cells_per_gpu": ens.ncell, "ms_per_step": ms,
                     "roofline": {"bound": "hbm", "kernel": "rconst_kernel<1>", "achieved": by / (ms * 1e-3) * 1e-9,
                                  "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                                  "peak_source": peak_src, "traffic": None,
                                  "note": "algorithmic bytes per cell = inputs as the host passes them (%d B, mostly the "
                                          "NSPEC-indexed exchange arrays) + NREACT*8 out" % (inb // ens.ncell)}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = min(ens.ncell, 20000)
        reps, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 3.0:
            ens.rconst(sl=slice(0, m))
            reps += 1
        dt = time.perf_counter() - t1
        res["rconst"]["cpu_baseline"] = {"value": m * reps / dt, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
                                         "sample": "%d x %d cells, host producer libmistra_rconst.so (OpenMP), %.1f s" % (reps, m, dt)}
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
