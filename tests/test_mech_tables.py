"""Consistency of the committed mechanism tables (mistra_b200/mech/*.json) -
host logic, no GPU.  The tables are the single source every generated line of C
and CUDA comes from, so their internal consistency is the first parity gate."""
import numpy as np
import pytest

from mistra_b200.mechgen import mech as mechmod

SIZES = {"gas": (102, 3, 331, 1110), "aer": (257, 5, 979, 6579), "tot": (417, 7, 1627, 13503)}
# SURVEY.md 8d flop table (re-derived from the reference sources)
FLOPS = {"gas": (1793, 1857, 1212, 6844, 2118), "aer": (4969, 4770, 6836, 142672, 12901),
         "tot": (8133, 7684, 13920, 434930, 26589)}


@pytest.mark.parametrize("name", mechmod.MECH_NAMES)
def test_sizes_match_parameters_headers(name):
    m = mechmod.load(name)
    assert (m.nvar, m.nfix, m.nreact, m.lu_nonzero) == SIZES[name]
    assert len(m.spc_names) == m.nvar + m.nfix == m.nspec
    assert len(m.reactions) == m.nreact and len(m.vdot) == m.nvar and len(m.jvs) == m.lu_nonzero
    assert len(m.d["rconst"]) == m.nreact


@pytest.mark.parametrize("name", mechmod.MECH_NAMES)
def test_flop_counts_match_survey(name):
    f = mechmod.load(name).flops
    assert (f["fun"], f["jac"], f["prep"], f["decomp"], f["solve"]) == FLOPS[name]


@pytest.mark.parametrize("name", mechmod.MECH_NAMES)
def test_sparse_pattern_is_sorted_csr_with_diagonal(name):
    m = mechmod.load(name)
    assert m.crow[0] == 0 and m.crow[m.nvar] == m.lu_nonzero
    for k in range(m.nvar):
        cols = m.icol[m.crow[k]:m.crow[k + 1]]
        assert (np.diff(cols) > 0).all()
        assert m.icol[m.diag[k]] == k and m.crow[k] <= m.diag[k] < m.crow[k + 1]


@pytest.mark.parametrize("name", mechmod.MECH_NAMES)
def test_pattern_is_closed_under_elimination(name):
    """LU fill-in must already be part of the pattern: eliminating row k with pivot
    row j never creates an entry outside row k's stored columns."""
    m = mechmod.load(name)
    rows = [set(m.icol[m.crow[k]:m.crow[k + 1]].tolist()) for k in range(m.nvar)]
    for k in range(m.nvar):
        for kk in range(m.crow[k], m.diag[k]):
            j = int(m.icol[kk])
            upper_j = set(m.icol[m.diag[j] + 1:m.crow[j + 1]].tolist())
            assert upper_j <= rows[k], (k, j)


@pytest.mark.parametrize("name", mechmod.MECH_NAMES)
def test_jacobian_tables_are_the_derivative_of_the_rate_tables(name):
    """Every B(m) must be dA(i)/dV(j) for exactly one (i, j), and
    JVS(row, col) = sum_i S[row, i] * dA(i)/dV(col) with the Vdot coefficients."""
    m = mechmod.load(name)

    def key(facs):
        rct = [f[1] for f in facs if f[0] == "R"]
        assert len(rct) == 1
        num = 1
        for f in facs:
            if f[0] == "N":
                num *= int(f[1])
        rest = sorted((f[0], f[1]) for f in facs if f[0] in "VF")
        return rct[0], num, tuple(rest)

    # derivative table from the reactions
    deriv = {}
    for i, facs in enumerate(m.reactions):
        vs = [f[1] for f in facs if f[0] == "V"]
        for j in set(vs):
            mult = vs.count(j)
            rest = [f for f in facs if not (f[0] == "V" and f[1] == j)] + [["V", j]] * (mult - 1)
            deriv[(i, j)] = (key(rest)[0], mult, key(rest)[2])
    inv = {}
    for (i, j), kv in deriv.items():
        inv.setdefault(kv, []).append((i, j))
    bmap = {}
    for b, facs in m.B:
        kv = key(facs)
        cands = inv.get(kv)
        assert cands, ("B(%d) is not a derivative of any reaction" % (b + 1), facs)
        bmap[b] = cands
    # stoichiometric matrix with literal strings
    S = {}
    for s, terms in enumerate(m.vdot):
        for sign, coef, a in terms:
            S.setdefault((s, a), []).append((sign, coef))
    for nz, terms in enumerate(m.jvs):
        row, col = int(m.row_of[nz]), int(m.icol[nz])
        expect = []
        for a in range(m.nreact):
            if (row, a) in S and (a, col) in deriv:
                expect += [(sg, cf, a) for sg, cf in S[(row, a)]]
        got = []
        for sign, coef, b in terms:
            rs = [i for (i, j) in bmap[b] if j == col and (row, i) in S]
            assert rs, (nz, b)
            got.append((sign, coef, rs))
        assert len(got) == len(expect), (nz, got, expect)
        for (sg, cf, rs) in got:
            assert any((sg, cf, a) in expect for a in rs), (nz, sg, cf, rs, expect)


@pytest.mark.parametrize("name", mechmod.MECH_NAMES)
def test_vdot_terms_are_in_ascending_reaction_order(name):
    """KPP writes every aggregate sum in reaction order; the scatter-form device
    code relies on it to reproduce the reference's left-to-right summation."""
    m = mechmod.load(name)
    for terms in m.vdot:
        idx = [a for _, _, a in terms]
        assert idx == sorted(idx)
