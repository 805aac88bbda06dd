"""Update_RCONST_g/_a/_t (rows a3 / N1) against an INDEPENDENT evaluation of the reference's own Fortran statements:
tests/golden/rconst_reference_*.npz were written by tests/golden/make_rconst_reference.py, which reads the RCONST(i) = ...
statements of gas.f / aer.f / tot.f and translates the rate-law functions of kpp.f90:7127-8373 line by line into Python -
it shares no code with mechgen/extract.py, csrc/rate_laws.h or the generated rconst_*.inc that the product compiles.
CPU: the host producer libmistra_rconst.so; GPU: rconst_kernel<mech> (mistra_rconst_update_device)."""
import os

import numpy as np
import pytest

from mistra_b200 import rconst as rcm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MECHS = [(0, "gas"), (1, "aer"), (2, "tot")]
KEYS = ("yhenry", "yxkmt", "ykef", "ykeb", "yxkmtd", "yxeq", "ycw", "ycwd")


def load(name):
    g = np.load(os.path.join(GOLD, "rconst_reference_%s.npz" % name))
    kw = {k: g[k] for k in KEYS if k in g.files}
    return g, kw


def check(got, ref, tol):
    assert got.shape == ref.shape
    assert np.array_equal(got == 0.0, ref == 0.0), "different switches: zero pattern of RCONST differs"
    nz = ref != 0.0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    assert rel.max() <= tol, "max rel err %.3e at reaction %d" % (rel.max(), int(np.nonzero(nz)[1][rel.argmax()]) + 1)


@pytest.mark.parametrize("mech,name", MECHS)
def test_host_producer_matches_the_reference_statements(mech, name):
    g, kw = load(name)
    got = rcm.update_rconst(mech, g["cb1"], g["scal"], g["ph_rat"], g["conc"], f32_literals=1, nthreads=1, **kw)
    # both sides use libm's exp / log10 / pow and the reference's operation order: a few ulp
    check(got, g["rconst"], 1e-13)


@pytest.mark.parametrize("mech,name", MECHS)
def test_every_reaction_is_exercised(mech, name):
    """The fixture cells switch on halogen and iodine chemistry, both aqueous bins and the dry-aerosol uptake, so
    (almost) every rate expression is evaluated with a non-zero result somewhere."""
    g, _ = load(name)
    used = (g["rconst"] != 0.0).any(axis=0)
    assert used.mean() >= 0.97, "only %.1f%% of the reactions have a non-zero rate constant in the fixture" % (100 * used.mean())


@pytest.mark.gpu
@pytest.mark.parametrize("mech,name", MECHS)
def test_device_kernel_matches_the_reference_statements(cuda_device, kpp, mech, name):
    import torch
    g, kw = load(name)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
    out = rcm.update_rconst_device(mech, t(g["cb1"]), t(g["scal"]), t(g["ph_rat"]), t(g["conc"]), f32_literals=1,
                                   **{k: t(v) for k, v in kw.items()})
    torch.cuda.synchronize()
    # CUDA's exp / pow / log10 differ from libm by a few ulp per operation
    check(out.cpu().numpy(), g["rconst"], 1e-12)
