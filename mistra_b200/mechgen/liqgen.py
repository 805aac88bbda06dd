"""Per-layer tables of the liq_parm chain (the part without reductions):

    henry_a / henry_t      kpp.f90:1914 / 1676   Henry constants                 henry(NSPEC)
    v_mean_a / v_mean_t    kpp.f90:1472 / 1268   mean molecular speeds           vmean(NSPEC)
    st_coeff_a / st_coeff_t kpp.f90:857 / 664    accommodation coefficients      alpha(NSPEC)
    equil_co_a / equil_co_t kpp.f90:3162 / 2954  forward / backward equilibrium  xkef, xkeb(NSPEC,nkc)

Each routine is a loop over the layers whose body is a list of scalar formulas in T (and conv2, xgamma) - no
reductions, several hundred statements.  As for Update_RCONST_x, the statements are DATA here:

  * `python -m mistra_b200.mechgen.liqgen --extract` (authoring container only, reads /root/reference) turns the loop
    bodies into C statements - species indices resolved against the mechanism's Parameters.h, default-REAL literals
    wrapped in RL() so that the binary32 / binary64 reading is a run-time switch (SURVEY 8a trap 1) - and stores them in
    mistra_b200/mech/liq_tables.json;
  * `python -m mistra_b200.mechgen.liqgen` (every build) emits csrc/_gen/liq_tables.inc from the JSON: one
    __host__ __device__ function per routine and mechanism, called once per layer by csrc/liq_kernels.cu (device) and
    csrc/rconst_host.cpp (host producer).

tests/golden/make_liq_reference.py evaluates the same Fortran statements with a Python back end (typed REAL arithmetic)
and pins both.
"""
from __future__ import annotations

import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
JSON = os.path.join(PKG, "mech", "liq_tables.json")
REF = "/root/reference/src"
ROUTINES = ("henry", "v_mean", "st_coeff", "equil_co")
OUT_ARRAYS = ("henry", "vmean", "alpha", "xkef", "xkeb")
NUM = re.compile(r"(?<![\w.])(\d+\.\d*|\.\d+|\d+)(?:([de])([+-]?\d+))?(_dp)?(?![\w])")
DECL = re.compile(r"^(use|implicit|real|integer|logical|include|common|external|double|character|parameter|save|intent)\b")


def strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def logical_lines(lines):
    out, cur = [], ""
    for ln in lines:
        ln = strip_comment(ln)
        if not ln.strip():
            continue
        t = ln.strip()
        if t.startswith("&"):
            t = t[1:]
        if t.endswith("&"):
            cur += t[:-1] + " "
            continue
        out.append(cur + t)
        cur = ""
    return out


def cexpr(e, par):
    """Fortran expression -> C expression (lower case)."""
    e = e.lower()
    for a, b in ((".eq.", "=="), (".ne.", "!="), (".le.", "<="), (".lt.", "<"), (".ge.", ">="), (".gt.", ">"),
                 (".and.", " && "), (".or.", " || "), (".not.", " !")):
        e = e.replace(a, b)
    assert "**" not in e, e

    def lit(m):
        mant, ex, exv, dp = m.group(1), m.group(2), m.group(3), m.group(4)
        if ex is None and dp is None and "." not in mant:
            return mant
        txt = mant + ("e" + exv if ex else "")
        if txt.endswith("."):
            txt += "0"
        if txt.startswith("."):
            txt = "0" + txt
        if "." not in txt.split("e")[0]:
            txt = txt.replace("e", ".0e") if "e" in txt else txt + ".0"
        if ex == "d" or dp:
            return txt
        return "RL(%s)" % txt
    e = NUM.sub(lit, e)
    e = re.sub(r"\bind_(\w+)\b", lambda m: str(par["ind_" + m.group(1)]), e)
    e = re.sub(r"\bmin\s*\(", "fmin(", e)
    e = re.sub(r"\bmax\s*\(", "fmax(", e)
    e = re.sub(r"\bdble\s*\(", "(double)(", e)
    return e


def matching_paren(s, i):
    d = 0
    for k in range(i, len(s)):
        if s[k] == "(":
            d += 1
        elif s[k] == ")":
            d -= 1
            if d == 0:
                return k
    raise ValueError("unbalanced: " + s)


def parameters(mech):
    txt = open(os.path.join(REF, "%s_Parameters.h" % mech), errors="replace").read()
    return {k.lower(): int(v) for k, v in re.findall(r"PARAMETER\s*\(\s*(\w+)\s*=\s*(\d+)\s*\)", txt)}


def translate(name, mech):
    """C statements of the layer-loop bodies of subroutine <name>."""
    par = parameters(mech)
    lines = open(os.path.join(REF, "kpp.f90"), errors="replace").read().split("\n")
    a = next(i for i, l in enumerate(lines) if re.match(r"^\s*subroutine\s+%s\b" % name, l, re.I))
    b = next(i for i in range(a, len(lines)) if re.match(r"^\s*end\s+subroutine\s+%s\b" % name, lines[i], re.I))
    out, locs = [], set()
    stack = []
    ind = 0

    def emit(s):
        out.append("  " * ind + s)
    for s in logical_lines(lines[a + 1:b]):
        low = s.lower().strip()
        if DECL.match(low):
            pm = re.search(r"parameter\s*::\s*(\w+)\s*=\s*(.+)$", s, re.I)
            if pm:
                emit("const double %s = %s;" % (pm.group(1).lower(), cexpr(pm.group(2), par)))
            continue
        m = re.match(r"^do\s+(\w+)\s*=\s*([^,]+),\s*(.+)$", low)
        if m:
            v = m.group(1)
            if v == "k":
                stack.append(None)                     # the loop over the layers: one call per layer
            else:
                emit("for (int %s = %s; %s <= %s; ++%s) {" % (v, cexpr(m.group(2), par), v, cexpr(m.group(3), par), v))
                stack.append(v)
                ind += 1
            continue
        if low in ("enddo", "end do"):
            if stack.pop() is not None:
                ind -= 1
                emit("}")
            continue
        if re.match(r"^(else\s*if|elseif)\b", low):
            i = s.index("(")
            j = matching_paren(s, i)
            ind -= 1
            emit("} else if (%s) {" % cexpr(s[i + 1:j], par))
            ind += 1
            continue
        if low.startswith("if"):
            i = s.index("(")
            j = matching_paren(s, i)
            rest = s[j + 1:].strip()
            if rest.lower() == "then":
                emit("if (%s) {" % cexpr(s[i + 1:j], par))
                ind += 1
                continue
            raise ValueError("one-line IF in %s: %s" % (name, s))
        if low == "else":
            ind -= 1
            emit("} else {")
            ind += 1
            continue
        if low in ("endif", "end if"):
            ind -= 1
            emit("}")
            continue
        whole = re.match(r"^(\w+)\s*\(\s*:\s*,\s*:\s*\)\s*=\s*(.+)$", s)
        if whole:                                      # alpha(:,:) = 0.1_dp
            arr = whole.group(1).lower()
            assert arr in OUT_ARRAYS
            emit("for (int j_ = 1; j_ <= nspec; ++j_) %s(j_, 0) = %s;" % (arr, cexpr(whole.group(2), par)))
            continue
        part = re.match(r"^(\w+)\s*\(\s*:\s*,([^)]*)\)\s*=\s*(.+)$", s)
        if part:                                       # xkef(:,kc,k) = 0._dp
            arr = part.group(1).lower()
            assert arr in OUT_ARRAYS
            emit("for (int j_ = 1; j_ <= nspec; ++j_) %s(j_,%s) = %s;" % (arr, cexpr(part.group(2), par), cexpr(part.group(3), par)))
            continue
        m = re.match(r"^(\w+)\s*(\(([^=]*)\))?\s*=(?!=)\s*(.+)$", s)
        if not m:
            raise ValueError("cannot translate %r in %s" % (s, name))
        lhs, args, rhs = m.group(1).lower(), m.group(3), m.group(4)
        if args is not None and lhs not in OUT_ARRAYS:
            # statement function: f(a0,b0[,k]) = expression
            an = [x.strip().lower() for x in args.split(",")]
            emit("auto %s = [&](%s) { return %s; };" % (lhs, ", ".join(("int " if x == "k" else "double ") + x for x in an),
                                                        cexpr(rhs, par)))
        elif args is not None:
            emit("%s(%s) = %s;" % (lhs, cexpr(args, par), cexpr(rhs, par)))
        else:
            locs.add(lhs)
            emit("%s = %s;" % (lhs, cexpr(rhs, par)))
    assert not stack and ind == 0, (name, stack, ind)
    return {"locals": sorted(locs), "body": out}


def extract():
    tab = {"note": "generated by mistra_b200/mechgen/liqgen.py --extract from the layer-loop bodies of henry_x, v_mean_x, "
                   "st_coeff_x, equil_co_x (kpp.f90:664-2145, 2954-3363); C statements over the macros of csrc/liq_tables.h",
           "cal15": 4.1855, "gas_const": 8.3144743}
    txt = open(os.path.join(REF, "constants.f90"), errors="replace").read()
    for k in ("cal15", "gas_const"):
        v = float(re.search(r"parameter\s*::\s*%s\s*=\s*([0-9.]+)_dp" % k, txt).group(1))
        assert v == tab[k], (k, v)
    for mech in ("aer", "tot"):
        tab[mech] = {r: translate("%s_%s" % (r, mech[0]), mech) for r in ROUTINES}
        tab[mech]["nspec"] = parameters(mech)["nspec"]
    with open(JSON, "w") as f:
        json.dump(tab, f, indent=0)
    print("wrote", JSON, {m: {r: len(tab[m][r]["body"]) for r in ROUTINES} for m in ("aer", "tot")})


def emit():
    tab = json.load(open(JSON))
    L = ["// GENERATED by mistra_b200/mechgen/liqgen.py from mistra_b200/mech/liq_tables.json - do not edit.",
         "// Layer-loop bodies of henry_x, v_mean_x, st_coeff_x, equil_co_x (kpp.f90:664-2145, 2954-3363) as C statements;",
         "// the macros henry(i,k) ... xgamma(j,kc,k), tt(k), t(k), conv2(kc,k), a_n2o5(k,kc) and RL() come from liq_tables.h.",
         "#pragma once"]
    for mech in ("aer", "tot"):
        x = mech[0]
        for r in ROUTINES:
            t = tab[mech][r]
            L.append("LIQ_HD void liq_%s_%s(LiqLayer &L_)" % (r, x))
            L.append("{")
            L.append("  const int nspec = %d; (void)nspec;" % tab[mech]["nspec"])
            L.append("  const double cal = %r, r = %r; (void)cal; (void)r;" % (tab["cal15"], tab["gas_const"]))
            L.append("  const int lpjoyce14bc = L_.lpjoyce14bc, lpbuxmann15alph = L_.lpbuxmann15alph; (void)lpjoyce14bc; (void)lpbuxmann15alph;")
            L.append("  const int k = 0, nkc = %d; (void)k; (void)nkc;" % (2 if mech == "aer" else 4))
            if t["locals"]:
                L.append("  double %s;" % ", ".join("%s = 0.0" % v for v in t["locals"]))
            L += ["  " + s for s in t["body"]]
            L.append("}")
            L.append("")
    out = os.path.join(PKG, "csrc", "_gen", "liq_tables.inc")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    text = "\n".join(L) + "\n"
    if not (os.path.exists(out) and open(out).read() == text):
        with open(out, "w") as f:
            f.write(text)
    print("wrote", out, len(text) // 1024, "KiB")


if __name__ == "__main__":
    if "--extract" in sys.argv:
        extract()
    emit()
