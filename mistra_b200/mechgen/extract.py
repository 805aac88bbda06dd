"""Build-time mechanism extractor.

Reads the KPP-generated mechanism sources of the reference (``src/gas.f``,
``src/aer.f``, ``src/tot.f`` + ``*_Parameters.h``) and writes a neutral,
compact description of each mechanism (``mistra_b200/mech/<name>.json``):

* sizes (NVAR, NFIX, NREACT, LU_NONZERO), species names, equation names;
* reactions: for every ``A(i)`` the ordered factor list of ``Fun_x``
  (``gas.f:2063-2393``);
* the aggregation ``Vdot = S*A`` with its literal coefficients, term order
  preserved (``gas.f:2396-2498``);
* the ``B(m)`` partial-rate products and the ``JVS(nz)`` sums of ``Jac_SP_x``
  (``gas.f:2675-6100``);
* the LU sparsity tables ``LU_ICOL/LU_CROW/LU_DIAG`` (``gas.f:6728-6832``);
* the statement order of the unrolled ``KppSolve_x`` (``gas.f:6217-6636``),
  which is cross-checked against the CSR tables.

The JSON tables are mechanism *data* (which reaction consumes which species
with which stoichiometric number); every line of C and CUDA in this repository
is generated from them by the oracle emitter / ``emit_cuda.py`` or written by
hand.  The reference tree is needed only to (re)run this script; the committed
tables make the repository buildable where ``/root/reference`` is absent (the
GPU box).

Usage:  python -m mistra_b200.mechgen.extract [/root/reference] [outdir]
"""
from __future__ import annotations

import json
import os
import re
import sys

MECHS = {"gas": "g", "aer": "a", "tot": "t"}


# --------------------------------------------------------------------------
# fixed-form Fortran helpers
# --------------------------------------------------------------------------
def _logical_lines(lines):
    """Join fixed-form continuation lines (non-blank, non-zero column 6),
    drop comment lines (C/c/*/! in column 1, or only a '!' comment)."""
    out = []
    for raw in lines:
        line = raw.rstrip("\n")
        if not line.strip():
            continue
        if line[0] in "Cc*!":
            continue
        if line.lstrip().startswith("!"):
            continue
        if len(line) > 5 and line[5] not in " 0" and line[:5].strip() == "":
            # continuation
            out[-1] += line[6:].strip()
        else:
            out.append(line[6:].strip() if len(line) > 6 else "")
    return out


def _routine(lines, header_re):
    """Return the physical lines of the program unit whose first statement
    matches header_re, up to its END."""
    start = None
    for i, l in enumerate(lines):
        if start is None:
            if l[:1] in "Cc*!":
                continue
            if re.search(header_re, l):
                start = i
        else:
            if re.match(r"^\s+END\s*(!.*)?$", l) or re.match(r"^\s+END\s*$", l):
                return lines[start:i], start + 1
    raise ValueError("routine %s not found" % header_re)


_num = r"[0-9]+(?:\.[0-9]*)?(?:[dDeE][+-]?[0-9]+)?"


def _parse_product(expr):
    """'RCT(7)*V(90)*F(1)*F(1)' or 'RCT(5)*2*V(3)' -> factor list.
    Each factor is ['R', i] | ['V', i] | ['F', i] | ['N', 'literal'];
    indices become 0-based."""
    facs = []
    for tok in expr.split("*"):
        m = re.fullmatch(r"(RCT|V|F)\((\d+)\)", tok)
        if m:
            facs.append([{"RCT": "R", "V": "V", "F": "F"}[m.group(1)], int(m.group(2)) - 1])
        elif re.fullmatch(_num, tok):
            facs.append(["N", tok])
        else:
            raise ValueError("unparsed factor %r in %r" % (tok, expr))
    return facs


def _parse_sum(expr, sym):
    """'-A(50)+0.63*A(54)+2*A(3)' -> [[sign, coef_literal|None, idx0], ...]
    preserving the left-to-right order the reference evaluates in."""
    terms = []
    pos = 0
    s = expr
    if s == "0":
        return []
    pat = re.compile(r"([+-]?)(?:(" + _num + r")\*)?" + sym + r"\((\d+)\)")
    while pos < len(s):
        m = pat.match(s, pos)
        if not m:
            raise ValueError("unparsed sum %r at %d" % (s, pos))
        sign = -1 if m.group(1) == "-" else 1
        terms.append([sign, m.group(2), int(m.group(3)) - 1])
        pos = m.end()
    return terms


def _parse_int_data(ll, name):
    """Collect 'DATA( NAME(i), i = a, b ) / ... /' integer blocks."""
    vals = []
    for l in ll:
        m = re.match(r"DATA\(\s*" + name + r"\(i\),\s*i\s*=\s*(\d+),\s*(\d+)\s*\)\s*/(.*)/$", l)
        if m:
            chunk = [int(x) for x in m.group(3).replace(" ", "").split(",") if x]
            assert len(chunk) == int(m.group(2)) - int(m.group(1)) + 1, (name, m.group(1))
            vals += chunk
            continue
        m = re.match(r"DATA\s*" + name + r"\s*/(.*)/$", l)
        if m:
            vals += [int(x) for x in m.group(1).replace(" ", "").split(",") if x]
    return vals


def _parse_str_data(lines, name):
    """String DATA blocks (SPC_NAMES / EQN_NAMES). Works on physical lines
    because quotes may contain anything."""
    vals = []
    active = False
    for l in lines:
        if re.search(r"DATA\(\s*" + name + r"\(i\)", l):
            active = True
            continue
        if active:
            body = l[6:] if len(l) > 6 else ""
            vals += re.findall(r"'([^']*)'", body)
            if body.rstrip().endswith("/"):
                active = False
    return [v.strip() for v in vals]


_RC_FUNCS = {"farr", "farr_sp", "atk_3", "atk_3f", "shno3", "fbck", "fbckj", "fbck2", "sp_17",
             "sp_23", "fcn", "dms_add", "fdhetg", "fdheta", "fdhett", "farr2", "dmin2", "dmin3",
             "uplim", "uparm", "uplip", "uparp", "flsc4", "flsc5", "flsc6", "fliq_60",
             "fhet_da", "fhet_t", "fhet_dt"}
_RC_SCALARS = {"conv1", "xhal", "xiod", "xhet1", "xhet2", "xliq1", "xliq2", "xliq3", "xliq4",
               "cvv1", "cvv2", "cvv3", "cvv4"}
_RC_ARR1 = {"ph_rat", "ycw"}                       # (n)       -> NAME(n-1)
_RC_SPC1 = {"yhenry", "c", "fix"}                  # (ind_X)   -> NAME(idx0)
_RC_SPC2 = {"yxkmt", "ykef", "ykeb"}               # (ind_X,k) -> NAME(idx0,k-1)


def _rconst_expr(e, ind, nvar):
    """Fortran right-hand side of one RCONST(i) statement -> neutral expression
    string in C syntax over a rate context `cx`:  functions f(cx,...), scalars
    S_name, arrays PH_RAT(i) YCW(i) YHENRY(s) C_(s) FIX_(s) YXKMT(s,k) ..."""
    indl = {k.lower(): v for k, v in ind.items()}
    toks = re.findall(r"[0-9]+\.?[0-9]*(?:[dDeE][+-]?[0-9]+)?|\.[0-9]+(?:[dDeE][+-]?[0-9]+)?"
                      r"|[A-Za-z_][A-Za-z_0-9]*|\*\*|[-+*/(),]", e)
    assert "".join(toks) == e, (e, toks)
    out = []
    i = 0
    while i < len(toks):
        t = toks[i]
        tl = t.lower()
        if re.match(r"[0-9.]", t):
            if re.search(r"[dD]", t):
                out.append(t.lower().replace("d", "e"))
            elif re.fullmatch(r"[0-9]+", t):
                out.append(t)                       # integer literal (function argument)
            else:
                out.append(t)
        elif tl in _RC_FUNCS:
            assert toks[i + 1] == "("
            if toks[i + 2] == ")":
                out.append("%s(cx" % tl)
            else:
                out.append("%s(cx," % tl)
            i += 1
        elif tl in _RC_SCALARS:
            out.append("S_" + tl)
        elif tl in _RC_ARR1:
            assert toks[i + 1] == "(" and toks[i + 3] == ")"
            out.append("%s(%d)" % (tl.upper(), int(toks[i + 2]) - 1))
            i += 3
        elif tl in _RC_SPC1:
            assert toks[i + 1] == "(" and toks[i + 3] == ")"
            v = indl[toks[i + 2].lower()] - 1
            nm = {"c": "C_", "fix": "FIX_", "yhenry": "YHENRY"}[tl]
            out.append("%s(%d)" % (nm, v))
            i += 3
        elif tl in _RC_SPC2:
            assert toks[i + 1] == "(" and toks[i + 3] == "," and toks[i + 5] == ")"
            v = indl[toks[i + 2].lower()] - 1
            out.append("%s(%d,%d)" % (tl.upper(), v, int(toks[i + 4]) - 1))
            i += 5
        elif t in "-+*/(),":
            out.append(t)
        else:
            raise ValueError("unparsed RCONST token %r in %r" % (t, e))
        i += 1
    return "".join(out)


# --------------------------------------------------------------------------
def extract(ref_root, mech):
    sfx = MECHS[mech]
    src = os.path.join(ref_root, "src", mech + ".f")
    with open(src, "r", errors="replace") as f:
        lines = f.readlines()

    par = open(os.path.join(ref_root, "src", mech + "_Parameters.h"), errors="replace").read()

    def P(n):
        return int(re.search(r"PARAMETER\s*\(\s*" + n + r"\s*=\s*(\d+)\s*\)", par).group(1))

    nvar, nfix, nreact, nspec, lunz = P("NVAR"), P("NFIX"), P("NREACT"), P("NSPEC"), P("LU_NONZERO")
    ind = {m.group(1): int(m.group(2)) for m in
           re.finditer(r"PARAMETER\s*\(\s*(indf?_\w+)\s*=\s*(\d+)\s*\)", par)}

    out = {"name": mech, "suffix": sfx, "nvar": nvar, "nfix": nfix, "nreact": nreact,
           "nspec": nspec, "lu_nonzero": lunz, "source": {}}

    # ---- Fun ---------------------------------------------------------------
    body, ln = _routine(lines, r"SUBROUTINE\s+Fun_" + sfx + r"\s*\(")
    out["source"]["Fun"] = "src/%s.f:%d" % (mech, ln)
    ll = _logical_lines(body)
    reactions = [None] * nreact
    vdot = [None] * nvar
    for l in ll:
        s = l.replace(" ", "")
        m = re.fullmatch(r"A\((\d+)\)=(.*)", s)
        if m:
            reactions[int(m.group(1)) - 1] = _parse_product(m.group(2))
            continue
        m = re.fullmatch(r"Vdot\((\d+)\)=(.*)", s)
        if m:
            vdot[int(m.group(1)) - 1] = _parse_sum(m.group(2), "A")
    assert all(r is not None for r in reactions), "missing A()"
    assert all(v is not None for v in vdot), "missing Vdot()"
    for r in reactions:
        assert r[0][0] == "R"
    out["reactions"] = reactions
    out["vdot"] = vdot

    # ---- Jac_SP ------------------------------------------------------------
    body, ln = _routine(lines, r"SUBROUTINE\s+Jac_SP_" + sfx + r"\s*\(")
    out["source"]["Jac_SP"] = "src/%s.f:%d" % (mech, ln)
    ll = _logical_lines(body)
    bdim = None
    B = {}
    jvs = [None] * lunz
    for l in ll:
        s = l.replace(" ", "")
        m = re.fullmatch(r"REAL\*8B\((\d+)\)", s)
        if m:
            bdim = int(m.group(1))
            continue
        m = re.fullmatch(r"B\((\d+)\)=(.*)", s)
        if m:
            B[int(m.group(1)) - 1] = _parse_product(m.group(2))
            continue
        m = re.fullmatch(r"JVS\((\d+)\)=(.*)", s)
        if m:
            jvs[int(m.group(1)) - 1] = _parse_sum(m.group(2), "B")
    assert bdim is not None
    assert all(j is not None for j in jvs), "missing JVS()"
    out["bdim"] = bdim
    out["B"] = [[k, B[k]] for k in sorted(B)]
    out["jvs"] = jvs

    # ---- sparse tables -----------------------------------------------------
    body, ln = _routine(lines, r"BLOCK\s+DATA\s+JACOBIAN_SPARSE_DATA_" + sfx)
    out["source"]["sparse"] = "src/%s.f:%d" % (mech, ln)
    ll = _logical_lines(body)
    # continuation marker in DATA blocks is '*' in column 6 -> handled above
    icol = _parse_int_data(ll, "LU_ICOL_" + sfx)
    crow = _parse_int_data(ll, "LU_CROW_" + sfx)
    diag = _parse_int_data(ll, "LU_DIAG_" + sfx)
    assert len(icol) == lunz and len(crow) == nvar + 1 and len(diag) == nvar + 1, \
        (len(icol), len(crow), len(diag))
    out["lu_icol"] = [c - 1 for c in icol]
    out["lu_crow"] = [c - 1 for c in crow]
    out["lu_diag"] = [c - 1 for c in diag]

    # ---- names -------------------------------------------------------------
    body, ln = _routine(lines, r"BLOCK\s+DATA\s+MONITOR_DATA_" + sfx)
    spc = _parse_str_data(body, "SPC_NAMES")
    eqn = _parse_str_data(body, "EQN_NAMES")
    assert len(spc) == nspec, (len(spc), nspec)
    assert len(eqn) == nreact, (len(eqn), nreact)
    out["spc_names"] = spc
    out["eqn_names"] = eqn
    for k, v in ind.items():
        nm = k.split("_", 1)[1]
        if k.startswith("indf_"):
            assert spc[nvar + v - 1] == nm, (k, v, spc[nvar + v - 1])
        else:
            assert spc[v - 1] == nm, (k, v, spc[v - 1])

    # ---- Update_RCONST ------------------------------------------------------
    body, ln = _routine(lines, r"SUBROUTINE\s+Update_RCONST_" + sfx)
    out["source"]["Update_RCONST"] = "src/%s.f:%d" % (mech, ln)
    ll = _logical_lines(body)
    rc = [None] * nreact
    for l in ll:
        s = l.replace(" ", "")
        m = re.fullmatch(r"RCONST\((\d+)\)=(.*)", s)
        if m:
            rc[int(m.group(1)) - 1] = _rconst_expr(m.group(2), ind, nvar)
    assert all(r is not None for r in rc), "missing RCONST()"
    out["rconst"] = rc

    # ---- KppSolve statement order (cross-check only) -----------------------
    body, ln = _routine(lines, r"SUBROUTINE\s+KppSolve_" + sfx + r"\s*\(")
    out["source"]["KppSolve"] = "src/%s.f:%d" % (mech, ln)
    ll = _logical_lines(body)
    solve = []
    for l in ll:
        s = l.replace(" ", "")
        m = re.fullmatch(r"X\((\d+)\)=(.*)", s)
        if not m:
            continue
        row = int(m.group(1)) - 1
        rhs = m.group(2)
        md = re.fullmatch(r"\((.*)\)/\(JVS\((\d+)\)\)", rhs)
        if md:
            inner, dg = md.group(1), int(md.group(2)) - 1
        else:
            md2 = re.fullmatch(r"X\((\d+)\)/JVS\((\d+)\)", rhs)
            if md2:
                inner, dg = "X(%d)" % (row + 1), int(md2.group(2)) - 1
            else:
                inner, dg = rhs, None
        assert inner.startswith("X(%d)" % (row + 1))
        rest = inner[len("X(%d)" % (row + 1)):]
        terms = [[int(a) - 1, int(b) - 1] for a, b in
                 re.findall(r"-JVS\((\d+)\)\*X\((\d+)\)", rest)]
        assert "".join("-JVS(%d)*X(%d)" % (a + 1, b + 1) for a, b in terms) == rest, rest
        solve.append([row, terms, dg])
    _check_solve(out, solve)
    out["source"]["KppDecomp"] = "src/%s.f:%d" % (mech, _routine(
        lines, r"SUBROUTINE\s+KppDecomp_" + sfx + r"\s*\(")[1])
    _check_pattern(out)
    return out


def _check_solve(m, solve):
    """The unrolled KppSolve must equal: forward rows ascending, strictly-lower
    entries in storage order (rows without lower entries omitted); backward rows
    descending, strictly-upper entries in storage order, divide by diagonal."""
    nvar, crow, diag, icol = m["nvar"], m["lu_crow"], m["lu_diag"], m["lu_icol"]
    exp = []
    for k in range(nvar):
        t = [[kk, icol[kk]] for kk in range(crow[k], diag[k])]
        if t:
            exp.append([k, t, None])
    for k in range(nvar - 1, -1, -1):
        t = [[kk, icol[kk]] for kk in range(diag[k] + 1, crow[k + 1])]
        exp.append([k, t, diag[k]])
    assert solve == exp, "KppSolve statements deviate from CSR-order substitution"


def _check_pattern(m):
    nvar, crow, diag, icol = m["nvar"], m["lu_crow"], m["lu_diag"], m["lu_icol"]
    for k in range(nvar):
        cols = icol[crow[k]:crow[k + 1]]
        assert cols == sorted(cols) and len(set(cols)) == len(cols)
        assert icol[diag[k]] == k
    assert crow[nvar] == m["lu_nonzero"]


def main(argv):
    ref = argv[1] if len(argv) > 1 else "/root/reference"
    outdir = argv[2] if len(argv) > 2 else os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mech")
    os.makedirs(outdir, exist_ok=True)
    for mech in MECHS:
        m = extract(ref, mech)
        path = os.path.join(outdir, mech + ".json")
        with open(path, "w") as f:
            json.dump(m, f, separators=(",", ":"))
        print("%s: nvar=%d nfix=%d nreact=%d lu_nonzero=%d B=%d -> %s (%d bytes)" % (
            mech, m["nvar"], m["nfix"], m["nreact"], m["lu_nonzero"], len(m["B"]),
            path, os.path.getsize(path)))


if __name__ == "__main__":
    main(sys.argv)
