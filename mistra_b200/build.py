"""Build the native pieces in-tree (sm_100a only):

  mistra_b200/libmistra_kpp.so          product library (C-ABI of include/mistra_kpp.h)
  mistra_b200/libmistra_kpp_strict.so   same sources with -DKPP_STRICT -fmad=false: divisions and
                                        summation order exactly as the reference; used by the
                                        bit-parity test only

Sources are generated from mistra_b200/mech/*.json by mechgen.emit_cuda, compiled
with nvcc (cross-compiles without a GPU) and linked with -cudart shared-free
static runtime.  Re-running is incremental (content hashes).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", HOSTCXX]

UNITS = ["kpp_mech_g.cu", "kpp_mech_a.cu", "kpp_mech_t.cu", "kpp_onchip_g.cu", "kpp_onchip_a.cu", "kpp_api.cu", "bins_kernels.cu",
         "kon_kernels.cu", "konc_kernels.cu", "cwrc_kernels.cu", "fastkmt_kernels.cu", "difc_kernels.cu", "drive_kernels.cu", "rconst_kernels.cu", "liq_kernels.cu", "sed_kernels.cu", "driver_kernels.cu", "_gen/kpp_names.cpp",
         "_gen/onchip_tables_g_%(v)s.cpp", "_gen/onchip_tables_a_%(v)s.cpp"]
# per-unit flags: the condensation kernel keeps the reference's unfused arithmetic
UNIT_FLAGS = {"kon_kernels.cu": ["-fmad=false"], "konc_kernels.cu": ["-fmad=false"], "cwrc_kernels.cu": ["-fmad=false"], "fastkmt_kernels.cu": ["-fmad=false"], "difc_kernels.cu": ["-fmad=false"], "drive_kernels.cu": ["-fmad=false"],
              "rconst_kernels.cu": ["-fmad=false"], "liq_kernels.cu": ["-fmad=false"], "sed_kernels.cu": ["-fmad=false"], "driver_kernels.cu": ["-fmad=false"]}


def _hash(paths, extra):
    h = hashlib.sha256()
    h.update(repr(extra).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _deps(unit):
    deps = [os.path.join(CSRC, unit), os.path.join(CSRC, "kpp_batch.h"),
            os.path.join(ROOT, "include", "mistra_kpp.h")]
    if unit.startswith("bins_"):
        deps.append(os.path.join(ROOT, "include", "mistra_bins.h"))
    if unit.startswith("kon_"):
        deps.append(os.path.join(ROOT, "include", "mistra_kon.h"))
    if unit.startswith("konc_"):
        deps.append(os.path.join(ROOT, "include", "mistra_konc.h"))
    if unit.startswith("cwrc_"):
        deps += [os.path.join(ROOT, "include", "mistra_cwrc.h"), os.path.join(CSRC, "tma_bulk.h")]
    if unit.startswith("fastkmt_"):
        deps += [os.path.join(ROOT, "include", "mistra_fastkmt.h"), os.path.join(CSRC, "tma_bulk.h")]
    if unit.startswith("difc_"):
        deps.append(os.path.join(ROOT, "include", "mistra_difc.h"))
    if unit.startswith("drive_"):
        deps.append(os.path.join(ROOT, "include", "mistra_drive.h"))
    if unit.startswith("driver_"):
        deps.append(os.path.join(ROOT, "include", "mistra_driver.h"))
    if unit.startswith("sed_"):
        deps.append(os.path.join(ROOT, "include", "mistra_sed.h"))
    if unit.startswith("liq_"):
        deps += [os.path.join(ROOT, "include", "mistra_liq.h"), os.path.join(CSRC, "liq_tables.h"),
                 os.path.join(CSRC, "liq_host.inc"), os.path.join(CSRC, "_gen", "liq_tables.inc")]
    if unit.startswith("rconst_"):
        deps += [os.path.join(ROOT, "include", "mistra_rconst.h"), os.path.join(ROOT, "include", "mistra_rconst_cuda.h"),
                 os.path.join(CSRC, "rate_laws.h"), os.path.join(CSRC, "rconst_common.h")]
        deps += [os.path.join(CSRC, "_gen", "rconst_%s.inc" % x) for x in "gat"]
    if unit.startswith("kpp_mech_"):
        x = unit[len("kpp_mech_")]
        deps += [os.path.join(CSRC, "_gen", "mech_%s.cuh" % x), os.path.join(CSRC, "ros3_kernel.inc")]
    if unit.startswith("kpp_onchip_"):
        x = unit[len("kpp_onchip_")]
        deps += [os.path.join(CSRC, "_gen", "onchip_%s.cuh" % x), os.path.join(CSRC, "ros3_onchip.inc"),
                 os.path.join(CSRC, "kpp_onchip.h")]
    if unit == "kpp_api.cu":
        deps += [os.path.join(CSRC, "kpp_onchip.h"), os.path.join(ROOT, "include", "mistra_kpp_rates.h"),
                 os.path.join(ROOT, "include", "mistra_rconst_cuda.h"), os.path.join(ROOT, "include", "mistra_rconst.h")]
    return deps


def _compile(unit, flags, tag, verbose):
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, "%s.%s.o" % (unit.replace("/", "_"), tag))
    stamp = obj + ".sha"
    flags = list(flags) + [f for f in UNIT_FLAGS.get(unit, []) if f not in flags]
    hv = _hash(_deps(unit), (flags, NVCC))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == hv:
        return obj
    cmd = [NVCC] + ARCH + COMMON + flags + ["-c", os.path.join(CSRC, unit), "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (unit, r.stdout))
    if verbose and r.stdout.strip():
        print(r.stdout)
    with open(stamp, "w") as f:
        f.write(hv)
    return obj


def build(verbose=False, strict=True, ptxas_v=False):
    subprocess.check_call([sys.executable, "-m", "mistra_b200.mechgen.emit_cuda"], cwd=ROOT,
                          stdout=subprocess.DEVNULL)
    subprocess.check_call([sys.executable, "-m", "mistra_b200.mechgen.onchip"], cwd=ROOT,
                          stdout=subprocess.DEVNULL)
    subprocess.check_call([sys.executable, "-m", "mistra_b200.mechgen.liqgen"], cwd=ROOT,
                          stdout=subprocess.DEVNULL)
    variants = [("fast", [], "libmistra_kpp.so")]
    if strict:
        variants.append(("strict", ["-DKPP_STRICT", "-fmad=false"], "libmistra_kpp_strict.so"))
    jobs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for tag, flags, _ in variants:
            fl = flags + (["-Xptxas", "-v"] if ptxas_v else [])
            for u in UNITS:
                jobs.append((tag, ex.submit(_compile, u % {"v": tag}, fl, tag, verbose)))
        objs = {}
        for tag, j in jobs:
            objs.setdefault(tag, []).append(j.result())
    outs = []
    for tag, flags, soname in variants:
        so = os.path.join(HERE, soname)
        newest = max(os.path.getmtime(o) for o in objs[tag])
        if not os.path.exists(so) or os.path.getmtime(so) < newest:
            cmd = [NVCC] + ARCH + ["-shared", "-ccbin", HOSTCXX, "-cudart", "static", "-o", so] + objs[tag]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        outs.append(so)
    outs.append(build_rconst(verbose))
    outs.append(build_f77(verbose))
    outs.append(build_b1_host(verbose))
    return outs


def build_variant(tag, extra_flags, units=("kpp_mech_a.cu",), verbose=False, base="fast"):
    """Experiment build: libmistra_kpp_<tag>.so = the product (base="fast") or strict (base="strict") objects
    with `units` recompiled with extra nvcc flags (e.g. -DKPP_MIN_BLOCKS=3).  Select it at run time with
    MISTRA_KPP_LIB=libmistra_kpp_<tag>.so (mistra_b200/kpp.py)."""
    bflags = ["-DKPP_STRICT", "-fmad=false"] if base == "strict" else []
    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        jobs = [ex.submit(_compile, u % {"v": base}, bflags + (list(extra_flags) if u in units else []),
                          tag if u in units else base, verbose)
                for u in UNITS]
        objs = [j.result() for j in jobs]
    so = os.path.join(HERE, "libmistra_kpp_%s.so" % tag)
    subprocess.check_call([NVCC] + ARCH + ["-shared", "-ccbin", HOSTCXX, "-cudart", "static", "-o", so] + objs)
    return so


def build_f77(verbose=False):
    """Boundary B1 shims (include/mistra_kpp_f77.h) over libmistra_kpp.so."""
    so = os.path.join(HERE, "libmistra_kpp_f77.so")
    src = os.path.join(CSRC, "f77_shim.c")
    deps = [src, os.path.join(ROOT, "include", "mistra_kpp_f77.h"), os.path.join(ROOT, "include", "mistra_kpp.h")]
    hv = _hash(deps, HOSTCXX)
    stamp = os.path.join(OBJ, "f77.sha")
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == hv:
        return so
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-O2", "-fPIC", "-shared", "-o", so, src, "-L" + HERE, "-lmistra_kpp", "-ldl",
           "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(hv)
    return so


def build_b1_host(verbose=False):
    """tests/host/libb1_host.so: a stand-in for the Fortran host of boundary B1 (defines the COMMON blocks, calls the
    shims) - test / bench infrastructure, not part of the product."""
    src = os.path.join(ROOT, "tests", "host", "b1_host.c")
    so = os.path.join(ROOT, "tests", "host", "libb1_host.so")
    deps = [src, os.path.join(ROOT, "include", "mistra_kpp_f77.h")]
    hv = _hash(deps, HOSTCXX)
    stamp = os.path.join(OBJ, "b1_host.sha")
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == hv:
        return so
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-O2", "-fPIC", "-shared", "-o", so, src, "-L" + HERE, "-lmistra_kpp_f77",
           "-Wl,-rpath," + HERE]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(hv)
    return so


def build_rconst(verbose=False):
    """Host-side Update_RCONST_x producer (include/mistra_rconst.h), plain g++."""
    so = os.path.join(HERE, "libmistra_rconst.so")
    srcs = [os.path.join(CSRC, "rconst_host.cpp"), os.path.join(CSRC, "_gen", "kpp_names.cpp")]
    deps = srcs + [os.path.join(CSRC, "rate_laws.h"), os.path.join(CSRC, "rconst_common.h"),
                   os.path.join(CSRC, "liq_tables.h"), os.path.join(CSRC, "liq_host.inc"), os.path.join(CSRC, "_gen", "liq_tables.inc"),
                   os.path.join(ROOT, "include", "mistra_liq.h"),
                   os.path.join(ROOT, "include", "mistra_rconst.h")] + [
        os.path.join(CSRC, "_gen", "rconst_%s.inc" % x) for x in "gat"]
    hv = _hash(deps, HOSTCXX)
    stamp = os.path.join(OBJ, "rconst.sha")
    os.makedirs(OBJ, exist_ok=True)
    if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == hv:
        return so
    cmd = [HOSTCXX, "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", so] + srcs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(hv)
    return so


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "--variant":
    # python -m mistra_b200.build --variant mb3 -DKPP_MIN_BLOCKS=3 [--units kpp_mech_a.cu,kpp_mech_g.cu]
    units = ("kpp_mech_a.cu",)
    flags = []
    base = "fast"
    for a in sys.argv[3:]:
        if a.startswith("--units="):
            units = tuple(a.split("=", 1)[1].split(","))
        elif a.startswith("--base="):
            base = a.split("=", 1)[1]
        else:
            flags.append(a)
    print(build_variant(sys.argv[2], flags, units, verbose=True, base=base))
    sys.exit(0)

if __name__ == "__main__":
    print("\n".join(build(verbose=True, strict="--no-strict" not in sys.argv,
                          ptxas_v="--ptxas-v" in sys.argv)))
