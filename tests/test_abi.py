"""The C-ABI library on a CPU-only box: it loads, exports every symbol the headers
declare, answers the metadata queries, validates options like Rosenbrock_x, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mistra_b200.mechgen import mech as mechmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mistra_\w+)\s*\(", txt)))


@pytest.mark.parametrize("header,lib", [("mistra_kpp.h", "libmistra_kpp.so"),
                                        ("mistra_kpp.h", "libmistra_kpp_strict.so"),
                                        ("mistra_bins.h", "libmistra_kpp.so"),
                                        ("mistra_kon.h", "libmistra_kpp.so"),
                                        ("mistra_konc.h", "libmistra_kpp.so"),
                                        ("mistra_cwrc.h", "libmistra_kpp.so"),
                                        ("mistra_fastkmt.h", "libmistra_kpp.so"),
                                        ("mistra_difc.h", "libmistra_kpp.so"),
                                        ("mistra_drive.h", "libmistra_kpp.so"),
                                        ("mistra_rconst_cuda.h", "libmistra_kpp.so"),
                                        ("mistra_kpp_rates.h", "libmistra_kpp.so"),
                                        ("mistra_liq.h", "libmistra_kpp.so"),
                                        ("mistra_sed.h", "libmistra_kpp.so"),
                                        ("mistra_driver.h", "libmistra_kpp.so"),
                                        ("mistra_rconst.h", "libmistra_rconst.so")])
def test_library_exports_every_declared_symbol(kpp, header, lib):
    L = C.CDLL(os.path.join(ROOT, "mistra_b200", lib))
    names = declared_functions(header)
    assert names
    for n in names:
        assert hasattr(L, n), "%s does not export %s" % (lib, n)


def test_b1_shim_exports_the_fortran_entry_points(kpp):
    """Boundary B1: lower-case names with a trailing underscore (gfortran / ifort), plus the optional bind call; the
    library loads into a process without a Fortran host, and an unbound COMMON block is reported, not dereferenced."""
    L = C.CDLL(os.path.join(ROOT, "mistra_b200", "libmistra_kpp_f77.so"))
    for n in ("integrate_g_", "integrate_a_", "integrate_t_", "mistra_kpp_f77_bind"):
        assert hasattr(L, n)
    host = os.path.join(ROOT, "tests", "host", "libb1_host.so")
    assert os.path.exists(host), "tests/host/libb1_host.so is built by mistra_b200.build"
    H = C.CDLL(host)      # local scope: a global load would let the product library's symbols interpose other test harnesses
    for n in ("gdata_g_", "gdata_a_", "gdata_t_", "b1_integrate", "b1_latency_us"):
        assert hasattr(H, n)
    for name, sym in (("gas", "gdata_g_"), ("aer", "gdata_a_"), ("tot", "gdata_t_")):
        m = mechmod.load(name)
        doubles = (m.nvar + m.nfix) + m.nreact + 2 + 2 * m.nvar + 2      # gas_Global.h:28-58
        nxt = {"gdata_g_": 644, "gdata_a_": 1759, "gdata_t_": 2889}[sym]
        assert doubles == nxt


def test_query_and_species_names(kpp):
    for mi, name in enumerate(mechmod.MECH_NAMES):
        m = mechmod.load(name)
        assert kpp.query(mi) == (m.nvar, m.nfix, m.nreact, m.lu_nonzero)
        assert [kpp.spc_name(mi, i) for i in range(m.nspec)] == m.spc_names
        assert kpp.spc_name(mi, m.nspec) is None and kpp.spc_name(mi, -1) is None
    with pytest.raises(kpp.KppError):
        kpp.query(3)


def test_default_opts_are_those_of_INTEGRATE(kpp):
    o = kpp.default_opts()
    assert (o.rtol, o.atol, o.hstart) == (1e-3, 1e-25, 1e-3)          # gas.f:743-746
    assert (o.hmin, o.hmax, o.facmin, o.facmax, o.facrej, o.facsafe, o.max_steps) == (0,) * 7
    assert o.f32_literals == 1 and o.autonomous == 0


@pytest.mark.parametrize("field,value", [("rtol", 2.0), ("rtol", 1e-16), ("atol", 0.0), ("hmin", -1.0),
                                         ("hmax", -1.0), ("hstart", -1.0), ("facmin", -1.0),
                                         ("facsafe", -0.5), ("max_steps", -3)])
def test_bad_options_are_rejected_before_touching_the_device(kpp, field, value):
    m = mechmod.load("gas")
    o = kpp.default_opts(**{field: value})
    with pytest.raises(kpp.KppError, match="error -2"):
        kpp.integrate(0, np.ones((1, m.nreact)), np.ones((1, m.nfix)), np.ones((1, m.nvar)), opts=o)


def test_compute_without_gpu_fails_loudly(kpp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = mechmod.load("gas")
    with pytest.raises(kpp.KppError):
        kpp.integrate(0, np.ones((2, m.nreact)), np.ones((2, m.nfix)), np.ones((2, m.nvar)))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under mistra_b200/ may reference it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "mistra_b200")):
        if "build" in dp.split(os.sep) or "_gen" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".inc", ".cuh")):
                t = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"kpp_oracle|from oracle|import oracle|oracle/", t):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
