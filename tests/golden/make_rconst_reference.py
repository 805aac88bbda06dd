"""Independent evaluator of Update_RCONST_g/_a/_t: executes the REFERENCE'S OWN Fortran statements.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_rconst_reference.py        # writes tests/golden/rconst_reference_{gas,aer,tot}.npz

What it does, without touching mistra_b200/mechgen/extract.py, csrc/rate_laws.h or the generated rconst_*.inc:
  * reads the species indices from src/{gas,aer,tot}_Parameters.h (PARAMETER ( ind_X = n ));
  * reads the statements  RCONST(i) = (...)  of SUBROUTINE Update_RCONST_x from src/{gas,aer,tot}.f
    (gas.f:275-666, aer.f:304-1400, tot.f:1040-2805), joining fixed-form continuation lines;
  * translates every rate-law FUNCTION of src/kpp.f90 (7127-8373: farr, atk_3, fhet_t, uparp, ...) line by line into
    Python (assignments, IF / ELSE IF / ELSE / ENDIF, one-line IF, STOP) and the right-hand sides into Python
    expressions; COMMON variables and arrays become names / callables of the evaluation namespace;
  * default-REAL literals (300., 8.314, 1.e-06: no D exponent, no _dp) are binary32 values with Fortran's mixed-mode
    rules (class F32: REAL op REAL / INTEGER is rounded to binary32 - e.g. 10**(-6.16) in fcn -, REAL op DOUBLE is
    promoted), as under the reference's preferred compiler flags (SURVEY 8a trap 1, f32_literals = 1);
  * evaluates RCONST for a handful of cells of the seeded synthetic ensembles and stores inputs and results.

tests/test_rconst_reference.py compares the host producer (libmistra_rconst.so) and, on the GPU, rconst_kernel with
these fixtures: rows a3 / N1 are then pinned by an evaluator that shares no code with the product.
"""
import math
import os
import re
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"
NCELL = {"gas": 24, "aer": 16, "tot": 12}

DECL = re.compile(r"^(use|implicit|real|integer|logical|include|common|external|double|character|parameter|save|intent)\b")
NUM = re.compile(r"(?<![\w.])(\d+\.\d*|\.\d+|\d+)(?:([de])([+-]?\d+))?(_dp)?(?![\w])")


def _r32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


class F32(float):
    """A default-REAL (binary32) value with Fortran's mixed-mode rules: REAL op REAL / INTEGER stays REAL (the result
    is rounded to binary32), REAL op DOUBLE PRECISION is promoted to double (a plain Python float)."""

    def __new__(cls, x):
        return float.__new__(cls, _r32(float(x)))

    @staticmethod
    def _single(o):
        return isinstance(o, F32) or (isinstance(o, int) and not isinstance(o, bool))

    def _bin(self, o, f):
        if F32._single(o):
            return F32(f(float(self), float(o)))
        return f(float(self), float(o))

    def _rbin(self, o, f):
        if F32._single(o):
            return F32(f(float(o), float(self)))
        return f(float(o), float(self))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._rbin(o, lambda a, b: a + b)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._rbin(o, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._rbin(o, lambda a, b: a * b)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._rbin(o, lambda a, b: a / b)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._rbin(o, lambda a, b: a ** b)
    def __neg__(self): return F32(-float(self))
    def __pos__(self): return self


def f32(x):
    return F32(x)


def _intrinsic(fn):
    """exp, log10, ... of a REAL argument are REAL; of a double argument double."""
    def g(x):
        return F32(fn(float(x))) if isinstance(x, F32) else fn(x)
    return g


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


def expr(e):
    """Fortran expression -> Python expression (lower case; literals: D exponent / _dp = binary64, else binary32)."""
    e = e.lower()
    for a, b in ((".eq.", "=="), (".ne.", "!="), (".le.", "<="), (".lt.", "<"), (".ge.", ">="), (".gt.", ">"),
                 (".and.", " and "), (".or.", " or "), (".not.", " not ")):
        e = e.replace(a, b)

    def lit(m):
        mant, ex, exv, dp = m.group(1), m.group(2), m.group(3), m.group(4)
        if ex is None and dp is None and "." not in mant:
            return mant                                       # integer literal
        txt = mant + ("e" + exv if ex else "")
        if ex == "d" or dp:
            return repr(float(txt))
        return "f32(%r)" % float(txt)                         # default REAL
    e = NUM.sub(lit, e)
    e = re.sub(r"\bdble\s*\(", "float(", e)
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


def translate_function(name, args, body):
    """Free-form Fortran function body -> Python source."""
    out = ["def %s(%s):" % (name, ", ".join(args))]
    ind = 1
    ret = "_ret_" + name

    def emit(s):
        out.append("    " * ind + s)

    def stmt(s):
        s = s.strip()
        low = s.lower()
        if low.startswith("stop"):
            emit("raise RuntimeError(%r)" % s)
        elif low.startswith("print") or low.startswith("write"):
            emit("pass")                                      # diagnostics of the reference, no effect on the value
        elif low == "return":
            emit("return %s" % ret)
        else:
            sf = re.match(r"^([a-z_]\w*)\s*\(([\w\s,]*)\)\s*=(?!=)(.*)$", s, re.I)
            if sf:                                            # Fortran statement function: f(a,b) = expression
                emit("%s = lambda %s: %s" % (sf.group(1).lower(), sf.group(2).lower(), expr(sf.group(3))))
                return
            m = re.match(r"^([a-z_]\w*)\s*=(?!=)(.*)$", s, re.I)
            if not m:
                raise ValueError("cannot translate %r in %s" % (s, name))
            lhs = m.group(1).lower()
            emit("%s = %s" % (ret if lhs == name else lhs, expr(m.group(2))))
    emit("%s = None" % ret)
    for s in body:
        low = s.lower().strip()
        if not low:
            continue
        if DECL.match(low):
            pm = re.search(r"parameter.*::\s*(\w+)\s*=\s*(.+)$", s, re.I)       # a named constant with its value
            if pm:
                emit("%s = %s" % (pm.group(1).lower(), expr(pm.group(2))))
            continue
        if re.match(r"^(else\s*if|elseif)\b", low):
            i = s.index("(")
            j = matching_paren(s, i)
            ind -= 1
            emit("elif %s:" % expr(s[i + 1:j]))
            ind += 1
        elif low.startswith("if"):
            i = s.index("(")
            j = matching_paren(s, i)
            rest = s[j + 1:].strip()
            if rest.lower() == "then":
                emit("if %s:" % expr(s[i + 1:j]))
                ind += 1
            else:
                emit("if %s:" % expr(s[i + 1:j]))
                ind += 1
                stmt(rest)
                ind -= 1
        elif low == "else":
            ind -= 1
            emit("else:")
            ind += 1
        elif low in ("endif", "end if"):
            ind -= 1
        else:
            stmt(s)
    emit("return %s" % ret)
    return "\n".join(out)


def rate_functions():
    """All FUNCTIONs of kpp.f90 as Python source, keyed by lower-case name."""
    lines = open(os.path.join(REF, "kpp.f90"), errors="replace").read().split("\n")
    logical, cur = [], ""
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
        logical.append(cur + t)
        cur = ""
    funcs = {}
    i = 0
    while i < len(logical):
        m = re.match(r"^\s*(?:double precision\s+|real\s*\([^)]*\)\s*)?function\s+(\w+)\s*\(([^)]*)\)", logical[i], re.I)
        if m:
            name = m.group(1).lower()
            args = [a.strip().lower() for a in m.group(2).split(",") if a.strip()]
            j = i + 1
            body = []
            while not re.match(r"^\s*end\s*function", logical[j], re.I):
                body.append(logical[j])
                j += 1
            try:
                funcs[name] = translate_function(name, args, body)
            except ValueError:
                pass                                   # functions outside the rate-law block may use other constructs
            i = j
        i += 1
    return funcs


def parameters(mech):
    txt = open(os.path.join(REF, "%s_Parameters.h" % mech), errors="replace").read()
    return {k.lower(): int(v) for k, v in re.findall(r"PARAMETER\s*\(\s*(\w+)\s*=\s*(\d+)\s*\)", txt)}


def rconst_statements(mech):
    """[(i, python expression)] of SUBROUTINE Update_RCONST_<x> (fixed form, continuation = non-blank column 6)."""
    x = mech[0]
    lines = open(os.path.join(REF, "%s.f" % mech), errors="replace").read().split("\n")
    a = next(i for i, l in enumerate(lines) if re.match(r"^\s+SUBROUTINE Update_RCONST_%s\b" % x, l))
    b = next(i for i in range(a, len(lines)) if re.match(r"^\s+END\s*$", lines[i]))
    logical = []
    for ln in lines[a:b]:
        if not ln.strip() or ln[0] in "Cc!*":
            continue
        ln = strip_comment(ln)
        if len(ln) > 5 and ln[5] not in " 0" and ln[:5].strip() == "":
            logical[-1] += ln[6:].strip()
        else:
            logical.append(ln.strip())
    out = []
    for s in logical:
        m = re.match(r"^RCONST\((\d+)\)\s*=\s*(.*)$", s)
        if m:
            out.append((int(m.group(1)), expr(m.group(2))))
    return out


def evaluate(mech, inp, funcs):
    """RCONST [ncell][NREACT] from the reference's statements, one cell at a time."""
    par = parameters(mech)
    nspec, nreact = par["nspec"], par["nreact"]
    stm = rconst_statements(mech)
    assert sorted(i for i, _ in stm) == list(range(1, nreact + 1)), "statement coverage of Update_RCONST_%s" % mech[0]
    code = [(i, compile(e, "RCONST(%d)" % i, "eval")) for i, e in stm]
    ns = {"f32": f32, "exp": _intrinsic(math.exp), "log": _intrinsic(math.log), "log10": _intrinsic(math.log10),
          "sqrt": _intrinsic(math.sqrt), "min": min, "max": max, "abs": abs, "float": float, "int": int, "dmin1": min,
          "dmax1": max, "amin1": min, "amax1": max, "dexp": math.exp, "dlog": math.log, "dlog10": math.log10,
          "dsqrt": math.sqrt, "dabs": abs}
    ns.update(par)
    for name, src in funcs.items():
        exec(src, ns)
    ncell = inp["cb1"].shape[0]
    nkc = 4 if mech == "tot" else 2
    out = np.zeros((ncell, nreact))
    for c in range(ncell):
        cb1, sc = inp["cb1"][c], inp["scal"][c]
        ns.update(aircc=float(cb1[0]), te=float(cb1[1]), h2oppm=float(cb1[2]), pk=float(cb1[3]))
        names = ("conv1", "xhal", "xiod", "xhet1", "xhet2", "xliq1", "xliq2", "xliq3", "xliq4", "cvv1", "cvv2", "cvv3", "cvv4")
        ns.update({n: float(v) for n, v in zip(names, sc)})
        conc = inp["conc"][c]
        nvar = par["nvar"]
        ns["c"] = lambda i, conc=conc: float(conc[i - 1])
        ns["fix"] = lambda i, conc=conc, nvar=nvar: float(conc[nvar + i - 1])
        ns["ph_rat"] = lambda i, p=inp["ph_rat"][c]: float(p[i - 1])

        def arr1(a):
            return (lambda i, a=a: float(a[i - 1])) if a is not None else (lambda i: 0.0)

        def arr2(a):                                   # Fortran y(NSPEC, nk): stored here as [nk][NSPEC]
            return (lambda i, k, a=a: float(a[k - 1][i - 1])) if a is not None else (lambda i, k: 0.0)
        g = lambda k: inp[k][c] if inp.get(k) is not None else None   # noqa: E731
        ns["yhenry"], ns["yxeq"] = arr1(g("yhenry")), arr1(g("yxeq"))
        ns["yxkmt"], ns["ykef"], ns["ykeb"], ns["yxkmtd"] = arr2(g("yxkmt")), arr2(g("ykef")), arr2(g("ykeb")), arr2(g("yxkmtd"))
        ns["ycw"], ns["ycwd"] = arr1(g("ycw")), arr1(g("ycwd"))
        for i, co in code:
            out[c, i - 1] = eval(co, ns)
    return out


def sample_inputs(mech):
    from mistra_b200 import synthetic
    cls = {"gas": synthetic.GasEnsemble, "aer": synthetic.AerEnsemble, "tot": synthetic.TotEnsemble}[mech]
    ens = cls(2, seed=20261019)
    n = ens.ncell
    idx = np.unique(np.linspace(0, n - 1, NCELL[mech]).astype(int))
    # a few chemistry steps' worth of non-trivial concentrations without running the integrator: the ensemble's own
    # initial state (aqueous ions, H+ included) is what Update_RCONST_x reads
    inp = {"cb1": ens.cb1[idx], "scal": ens.scal[idx], "ph_rat": ens.ph_rat[idx], "conc": ens.conc()[idx]}
    for k in ("yhenry", "yxkmt", "ykef", "ykeb", "yxkmtd", "yxeq", "ycw", "ycwd"):
        a = getattr(ens, k, None)
        inp[k] = None if a is None else np.ascontiguousarray(a[idx])
    return inp


def main():
    funcs = rate_functions()
    print("translated %d functions of kpp.f90" % len(funcs))
    for mech in ("gas", "aer", "tot"):
        inp = sample_inputs(mech)
        rc = evaluate(mech, inp, funcs)
        assert np.isfinite(rc).all()
        out = os.path.join(ROOT, "tests", "golden", "rconst_reference_%s.npz" % mech)
        np.savez_compressed(out, rconst=rc, **{k: v for k, v in inp.items() if v is not None})
        print("%s: %d cells x %d reactions, %d non-zero -> %s" % (mech, rc.shape[0], rc.shape[1], int((rc != 0).sum()), out))


if __name__ == "__main__":
    main()
