"""Independent evaluation of henry_x, v_mean_x, st_coeff_x, equil_co_x (and a_n2o5) - row N2, the per-layer tables of the
liq_parm chain - by executing the reference's own Fortran statements with a Python back end (authoring container only):

    python tests/golden/make_liq_reference.py      # writes tests/golden/liq_reference_{aer,tot}.npz

Front end: the subroutines are read from /root/reference/src/kpp.f90 (664-2145, 2954-3363, 8377-8439), comments stripped,
continuation lines joined.  Back end: every statement becomes Python - the loop over the layers is dropped (one layer at
a time), DO kc / DO j loops, IF blocks, statement functions, whole-array assignments, array elements as indexing - and is
executed with the typed REAL arithmetic of make_rconst_reference.py (default-REAL literals are binary32 values, promoted
on contact with a double).  The product's generator (mistra_b200/mechgen/liqgen.py) emits C from the same text with a
different back end; tests/test_liq_reference.py holds the host build and the CUDA kernel to these fixtures.
"""
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_rconst_reference import REF, _intrinsic, expr, f32, matching_paren, parameters, strip_comment  # noqa: E402

ARR = ("henry", "vmean", "alpha", "xkef", "xkeb")
DECL = re.compile(r"^(use|implicit|real|integer|logical|include|common|external|double|character|parameter|save|intent)\b")
J6 = 55


class Out:
    """One layer of an output array indexed like the Fortran array (species[, bin], layer)."""

    def __init__(self, nspec, nk=None):
        self.a = np.zeros((nk, nspec) if nk else nspec)
        self.nk = nk

    def __getitem__(self, idx):
        return float(self.a[idx[1] - 1, idx[0] - 1]) if self.nk else float(self.a[idx[0] - 1])

    def __setitem__(self, idx, v):
        if self.nk:
            self.a[idx[1] - 1, idx[0] - 1] = float(v)
        else:
            self.a[idx[0] - 1] = float(v)


def index_arrays(e):
    """name(args) -> name[args] for the output arrays (matching parentheses)."""
    out = ""
    i = 0
    pat = re.compile(r"\b(%s)\s*\(" % "|".join(ARR))
    while True:
        m = pat.search(e, i)
        if not m:
            return out + e[i:]
        j = matching_paren(e, m.end() - 1)
        out += e[i:m.start()] + m.group(1) + "[" + index_arrays(e[m.end():j]) + "]"
        i = j + 1


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


def translate(name):
    lines = open(os.path.join(REF, "kpp.f90"), errors="replace").read().split("\n")
    a = next(i for i, l in enumerate(lines) if re.match(r"^\s*subroutine\s+%s\b" % name, l, re.I))
    b = next(i for i in range(a, len(lines)) if re.match(r"^\s*end\s+subroutine\s+%s\b" % name, lines[i], re.I))
    out = ["def %s():" % name, "    k = 0"]
    ind = 1
    stack = []

    def emit(s):
        out.append("    " * ind + s)

    def px(e):
        return index_arrays(expr(e))
    for s in logical_lines(lines[a + 1:b]):
        low = s.lower().strip()
        if DECL.match(low):
            pm = re.search(r"parameter\s*::\s*(\w+)\s*=\s*(.+)$", s, re.I)
            if pm:
                emit("%s = %s" % (pm.group(1).lower(), px(pm.group(2))))
            continue
        m = re.match(r"^do\s+(\w+)\s*=\s*([^,]+),\s*(.+)$", low)
        if m:
            if m.group(1) == "k":
                stack.append(None)
            else:
                emit("for %s in range(%s, (%s) + 1):" % (m.group(1), px(m.group(2)), px(m.group(3))))
                stack.append(m.group(1))
                ind += 1
            continue
        if low in ("enddo", "end do"):
            if stack.pop() is not None:
                ind -= 1
            continue
        if re.match(r"^(else\s*if|elseif)\b", low):
            i = s.index("(")
            j = matching_paren(s, i)
            ind -= 1
            emit("elif %s:" % px(s[i + 1:j]))
            ind += 1
            continue
        if low.startswith("if"):
            i = s.index("(")
            j = matching_paren(s, i)
            assert s[j + 1:].strip().lower() == "then", s
            emit("if %s:" % px(s[i + 1:j]))
            ind += 1
            continue
        if low == "else":
            ind -= 1
            emit("else:")
            ind += 1
            emit("pass")
            continue
        if low in ("endif", "end if"):
            ind -= 1
            continue
        whole = re.match(r"^(\w+)\s*\(\s*:\s*,\s*:\s*\)\s*=\s*(.+)$", s)
        if whole:
            emit("%s.a[...] = float(%s)" % (whole.group(1).lower(), px(whole.group(2))))
            continue
        part = re.match(r"^(\w+)\s*\(\s*:\s*,\s*(\w+)\s*,[^)]*\)\s*=\s*(.+)$", s)
        if part:
            emit("%s.a[%s - 1, :] = float(%s)" % (part.group(1).lower(), part.group(2).lower(), px(part.group(3))))
            continue
        m = re.match(r"^(\w+)\s*(\(([^=]*)\))?\s*=(?!=)\s*(.+)$", s)
        assert m, s
        lhs, args, rhs = m.group(1).lower(), m.group(3), m.group(4)
        if args is not None and lhs not in ARR:
            emit("%s = lambda %s: %s" % (lhs, args.lower(), px(rhs)))
        elif args is not None:
            emit("%s[%s] = %s" % (lhs, px(args), px(rhs)))
        else:
            emit("%s = %s" % (lhs, px(rhs)))
        if stack and stack[-1] is None and False:
            pass
    # IF blocks whose body is empty in the Fortran text (else branches with comments only)
    src = "\n".join(out)
    src = re.sub(r"(\n(\s*)if [^\n]*:\n)(?=\2(else:|elif ))", r"\1\2    pass\n", src)
    return src


def a_n2o5_factory(cw, cm, s13, s14):
    def a_n2o5(k, kc):
        xno3m = xclm = xh2o = 0.0
        denom = 1.0
        if cw[kc - 1] > 0.0:
            xno3m = s13[kc - 1] / cw[kc - 1] * 1e-3
            xclm = s14[kc - 1] / cw[kc - 1] * 1e-3
        if cm[0] > 0.0 and cw[0] > 0.0:
            xh2o = 55.55 * (cm[0] / cw[0])
        xk2f = 1.15e6 - 1.15e6 * math.exp(-0.13 * xh2o)
        if xno3m > 0.0:
            denom = 1.0 + 6.e-2 * xh2o / xno3m + 29.0 * xclm / xno3m
        return 3.2e-8 * xk2f * (1.0 - (1.0 / denom))
    return a_n2o5


def main():
    for mech, nkc in (("aer", 2), ("tot", 4)):
        par = parameters(mech)
        nspec = par["nspec"]
        srcs = [translate("%s_%s" % (r, mech[0])) for r in ("henry", "v_mean", "st_coeff", "equil_co")]
        r = np.random.default_rng(31 + nkc)
        ncell = 10
        t = r.uniform(235.0, 300.0, ncell)
        cw = 10.0 ** r.uniform(-12, -7, (ncell, nkc))
        cw[r.uniform(size=(ncell, nkc)) < 0.2] = 0.0                  # bins without liquid water
        conv2 = np.where(cw > 0, 1.0 / (1000.0 * np.where(cw > 0, cw, 1.0)), 0.0)
        cm = cw * r.uniform(0.5, 1.0, (ncell, nkc))
        xgamma = r.uniform(0.3, 1.5, (ncell, nkc, J6))
        s1314 = 10.0 ** r.uniform(-12, -8, (ncell, nkc, 2))
        res = {}
        for flags in ((0, 0), (1, 1)):
            H = np.zeros((ncell, nspec)); V = np.zeros((ncell, nspec)); A = np.zeros((ncell, nspec))
            KF = np.zeros((ncell, nkc, nspec)); KB = np.zeros((ncell, nkc, nspec))
            for c in range(ncell):
                ns = {"f32": f32, "exp": _intrinsic(math.exp), "sqrt": _intrinsic(math.sqrt), "log": _intrinsic(math.log),
                      "min": min, "max": max, "nspec": nspec, "nkc": nkc, "cal": 4.1855, "r": 8.3144743,
                      "lpjoyce14bc": bool(flags[0]), "lpbuxmann15alph": bool(flags[1]), "nf": 1, "nmaxf": 1}
                ns.update(par)
                ns["henry"], ns["vmean"], ns["alpha"] = Out(nspec), Out(nspec), Out(nspec)
                ns["xkef"], ns["xkeb"] = Out(nspec, nkc), Out(nspec, nkc)
                ns["tt"] = ns["t"] = lambda k, v=float(t[c]): v
                ns["conv2"] = lambda kc, k, a=conv2[c]: float(a[kc - 1])
                ns["xgamma"] = lambda j, kc, k, a=xgamma[c]: float(a[kc - 1, j - 1])
                ns["a_n2o5"] = a_n2o5_factory(cw[c], cm[c], s1314[c, :, 0], s1314[c, :, 1])
                for nm, src in zip(("henry", "v_mean", "st_coeff", "equil_co"), srcs):
                    exec(src, ns)
                    ns["%s_%s" % (nm, mech[0])]()
                H[c], V[c], A[c] = ns["henry"].a, ns["vmean"].a, ns["alpha"].a
                KF[c], KB[c] = ns["xkef"].a, ns["xkeb"].a
            res[flags] = (H, V, A, KF, KB)
        out = os.path.join(HERE, "liq_reference_%s.npz" % mech)
        np.savez_compressed(out, t=t, cw=cw, cm=cm, conv2=conv2, xgamma=xgamma, sion1_13_14=s1314,
                            **{"%s_%d%d" % (n, f[0], f[1]): v for f, vals in res.items()
                               for n, v in zip(("henry", "vmean", "alpha", "xkef", "xkeb"), vals)})
        H = res[(0, 0)][0]
        print("%s: %d layers, henry non-zero %d, alpha != 0.1 in %d species, xkef non-zero %d -> %s"
              % (mech, ncell, int((H != 0).any(axis=0).sum()), int((res[(0, 0)][2] != 0.1).any(axis=0).sum()),
                 int((res[(0, 0)][3] != 0).any(axis=(0, 1)).sum()), out))


if __name__ == "__main__":
    main()
