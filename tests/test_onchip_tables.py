"""The instruction streams of the on-chip Ros3 kernels (mistra_b200/mechgen/onchip.py), executed on
the CPU by an emulator that follows the kernel's interpreters op by op (IEEE double, no fused
multiply-add), against the CPU oracle's building blocks:

* strict plan (what libmistra_kpp_strict.so runs): Fun, Jac_SP + matrix preparation, the elimination
  (wave-scheduled head, head pivots on the register tail, right-looking dense tail) and both
  triangular sweeps are bit-identical to the reference formulation;
* fast plan (libmistra_kpp.so: reciprocal pivots, reordered backward sweep, split forward sums) agrees
  to a few ulps.

Also checks the structural invariants the kernel relies on: block barriers at the same position in
every warp's stream, no two operations of an elimination wave touching the same slot, the shared
memory budget of five (aer) / fifteen (gas) cell slots per block."""
import numpy as np
import pytest

from mistra_b200.mechgen import mech as mechmod
from mistra_b200.mechgen import onchip
from tests import util

GAMMA1 = 0.43586652150845899941601945119356
CASES = [(0, "gas"), (1, "aer")]


@pytest.fixture(scope="module", params=CASES, ids=[c[1] for c in CASES])
def plans(request):
    mi, name = request.param
    m = mechmod.load(name)
    T = onchip.TAIL[name]
    return mi, name, m, onchip.Plan(m, T, strict=True), onchip.Plan(m, T, strict=False)


def _cell(name, seed):
    var, fix, rc = util.random_cells(name, 1, seed)
    return var[0], fix[0], rc[0]


def test_streams_against_oracle_blocks(plans, oracle):
    mi, name, m, ps, pf = plans
    for c in range(2):
        V, F, R = _cell(name, 700 + 10 * mi + c)
        ghinv = 1.0 / (10.0 ** (-2 - c) * GAMMA1)
        fo = oracle.fun(mi, V, F, R, f32=1)
        J = -oracle.jac(mi, V, F, R, f32=1)
        J[m.diag[:m.nvar]] += ghinv
        LU, ier = oracle.decomp(mi, J)
        assert ier == 0
        X = np.random.default_rng(5 + c).normal(size=m.nvar)
        Xo = oracle.solve(mi, LU, X)
        for p in (ps, pf):
            S = p.new_smem()
            S[p.O_Y:p.O_Y + p.n] = V
            p.emu_set_consts(S, F, 1)
            p.emu_fun(S, p.O_Y, p.O_K1, R)
            assert np.array_equal(S[p.O_K1:p.O_K1 + p.n], fo)                  # Fun_x
            a, sing = p.emu_jacprep(S, R, ghinv)
            assert not sing
            assert np.array_equal(p.emu_to_csr(S, a), J)                       # Jac_SP_x + ros_PrepareMatrix
            p.emu_decomp(S, a)
            if p.strict:
                assert np.array_equal(p.emu_to_csr(S, a), LU)                  # KppDecomp_x, to the last bit
            S[p.O_K1:p.O_K1 + p.n] = X
            p.emu_solve(S, a, p.O_K1)
            Xs = S[p.O_K1:p.O_K1 + p.n]
            if p.strict:
                assert np.array_equal(Xs, Xo)                                   # KppSolve_x, to the last bit
            else:
                assert np.max(np.abs(Xs - Xo) / np.maximum(np.abs(Xo), 1e-300)) < 1e-12


def test_second_literal_variant(plans, oracle):
    """-r8 builds of the reference read the default-REAL literals as binary64 (SURVEY 8a trap 1)."""
    mi, name, m, ps, _ = plans
    V, F, R = _cell(name, 811 + mi)
    S = ps.new_smem()
    S[ps.O_Y:ps.O_Y + ps.n] = V
    ps.emu_set_consts(S, F, 0)
    ps.emu_fun(S, ps.O_Y, ps.O_K1, R)
    assert np.array_equal(S[ps.O_K1:ps.O_K1 + ps.n], oracle.fun(mi, V, F, R, f32=0))


def test_singular_diagonal_is_reported(plans):
    """KppDecomp's test JVS(LU_DIAG(k)).EQ.0 (gas.f:6156) on the prepared matrix: a cell with J = 0 and
    1/(H gamma) = 0 trips it, in the head (shared memory) and in the tail (registers) alike."""
    mi, name, m, ps, _ = plans
    V, F, R = _cell(name, 5)
    S = ps.new_smem()
    S[ps.O_Y:ps.O_Y + ps.n] = 0.0
    ps.emu_set_consts(S, F, 1)
    a, sing = ps.emu_jacprep(S, R * 0.0, 0.0)
    assert sing


def test_elimination_waves_have_no_hazards(plans):
    mi, name, m, ps, pf = plans
    for p in (ps, pf):
        for ops in p.hop_waves:
            written = [o[1] for o in ops]
            assert len(set(written)) == len(written)
            reads = set()
            for fl, d, c, a, b in ops:
                reads |= {c, a, b} - {d}
            assert not (reads & set(written) - {p.ZERO})


def test_block_barriers_line_up(plans):
    mi, name, m, ps, pf = plans
    for p in (ps, pf):
        for nm in ("fwd", "bwd"):
            P = p.unpack16(getattr(p, nm + "_stream"), getattr(p, nm + "_nchunk"))
            counts = []
            for w in range(p.W):
                ws = P[32 * w]
                flags = [int(ws[i + 1]) & 3 for i in range(0, len(ws), 8)]
                counts.append(sum(1 for f in flags if f & onchip.F_CSYNC))
                for lane in range(1, 32):
                    wl = P[32 * w + lane]
                    assert [int(wl[i + 1]) & 3 for i in range(0, len(wl), 8)] == flags
            assert len(set(counts)) == 1


def test_shared_memory_budget(plans):
    """One persistent block per SM: the cell slots + the 8-stage stream pipeline + barriers fit in 227 KiB."""
    mi, name, m, ps, pf = plans
    slots = {"gas": 15, "aer": 5}[name]
    for p in (ps, pf):
        assert slots * p.smem_doubles * 8 + 8 * 1024 * p.W + 2 * 8 * 8 + 2 * slots * 4 + 16 <= 232448
        assert p.smem_doubles == ps.smem_doubles and p.smem_doubles % 2 == 0
