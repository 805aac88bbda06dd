"""Independent evaluation of sedp, sedl, sedc (species loop), advsed0, advsed1 and vterm - row N4, the gravitational
settling operators - by executing the reference's own Fortran statements with a Python back end (authoring container
only, needs /root/reference):

    python tests/golden/make_sed_reference.py      # writes tests/golden/sed_reference.npz

Front end: the routines are read from /root/reference/src/str.f90 (sedp 2257-2411, sedc 2567-2596, sedl 2627-2787,
vterm 2793-2864, advsed0 5522-5579, advsed1 5585-5691), comments stripped, continuation lines joined.  Back end: every
executable statement becomes one Python statement - DO / DO WHILE / IF blocks, one-line IF, CALL, EXIT, assignments;
arrays are 1-based objects (read by call syntax exactly as the Fortran text writes them, written through .set), local
arrays are created from the routine's own declarations, COMMON variables live in the evaluation namespace.  Literals
follow make_rconst_reference.py (default-REAL literals are binary32 values promoted on contact with a double); x**3 is
the repeated product gfortran emits.  The model sizes (nf, nka, nkt, j2, j6, nkc) are names of that namespace, so the
fixtures use a small grid.  Shares no code with oracle/sed_oracle.c or mistra_b200/csrc/sed_kernels.cu;
tests/test_sed_oracle.py holds the oracle to these fixtures, tests/test_gpu_sed.py the CUDA kernels.
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
from make_rconst_reference import REF, _intrinsic, expr, matching_paren, strip_comment  # noqa: E402

DECL = re.compile(r"^(use|implicit|real|integer|logical|include|common|external|double|character|parameter|save|intent)\b")


class FA:
    """Fortran array: bounds per dimension, column-major meaning irrelevant here (element access only)."""

    def __init__(self, *bounds, data=None):
        self.lo = [b[0] if isinstance(b, tuple) else 1 for b in bounds]
        shape = [b[1] - b[0] + 1 if isinstance(b, tuple) else b for b in bounds]
        self.a = np.zeros(shape) if data is None else data
        assert list(self.a.shape) == shape, (self.a.shape, shape)

    def _ix(self, idx):
        assert len(idx) == len(self.lo)
        t = tuple(int(i) - l for i, l in zip(idx, self.lo))
        assert all(0 <= i < s for i, s in zip(t, self.a.shape)), ("subscript out of bounds", idx, self.a.shape)
        return t

    def __call__(self, *idx):
        v = self.a[self._ix(idx)]
        return int(v) if self.a.dtype.kind == "i" else float(v)

    def set(self, idx, v):
        self.a[self._ix(idx)] = float(v)


def frange(a, b, s=1):
    return range(a, b + (1 if s > 0 else -1), s)


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


def split_top(s):
    """split at top-level commas"""
    out, d, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            d += 1
        elif ch == ")":
            d -= 1
        if ch == "," and d == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return [x.strip() for x in out]


def pyexpr(e):
    e = re.sub(r"\b(\w+)\s*\*\*\s*3\b", r"(\1*\1*\1)", e)         # gfortran: x**3 -> x*x*x
    e = expr(e)
    return e.replace(".true.", " True ").replace(".false.", " False ")


def translate(first, last, name=None, args=(), arrays=(), result=None, commons=(), fname="str.f90"):
    """Lines first..last (1-based, inclusive) of a reference source -> Python source.  name=None: module-level
    statements."""
    text = open(os.path.join(REF, fname), errors="replace").read().split("\n")[first - 1:last]
    body = logical_lines(text)
    out, ind = [], 0
    known = set(arrays)

    def emit(s):
        out.append("    " * ind + s)

    if name:
        emit("def %s(%s):" % (name, ", ".join(args)))
        ind = 1
        if commons:
            emit("global " + ", ".join(commons))
        if result:
            emit("_ret = None")

    def stmt(s):
        low = s.lower().strip()
        if low.startswith("call "):
            emit(pyexpr(s.strip()[5:]))
        elif low == "exit":
            emit("break")
        elif low.startswith("print") or low.startswith("write"):
            emit("pass")
        else:
            m = re.match(r"^([a-z_]\w*)\s*(\(.*?\))?\s*=(?!=)(.*)$", s.strip(), re.I)
            if not m:
                raise ValueError("cannot translate %r" % s)
            lhs = m.group(1).lower()
            if m.group(2):
                # the subscript list ends at the parenthesis that matches the first one
                st = s.strip()
                i = st.index("(")
                j = matching_paren(st, i)
                rhs = st[j + 1:].lstrip()
                assert rhs.startswith("="), s
                assert lhs in known, ("assignment to an undeclared array", s)
                if st[i + 1:j].strip() == ":":                           # whole-array assignment  a(:) = scalar
                    emit("%s.a[...] = %s" % (lhs, pyexpr(rhs[1:])))
                    return
                emit("%s.set((%s,), %s)" % (lhs, pyexpr(st[i + 1:j]), pyexpr(rhs[1:])))
            else:
                emit("%s = %s" % ("_ret" if lhs == result else lhs, pyexpr(m.group(3))))

    for s in body:
        low = s.lower().strip()
        if re.match(r"^(subroutine|function|end subroutine|end function)\b", low):
            continue
        if DECL.match(low):
            pm = re.search(r"parameter\s*::\s*(\w+)\s*=\s*(.+)$", s, re.I)
            if pm:
                emit("%s = %s" % (pm.group(1).lower(), pyexpr(pm.group(2))))
            elif re.match(r"^real\b", low) and "::" in s and "intent" not in low:
                for item in split_top(s.split("::", 1)[1]):          # local arrays: c(nf), a0(2:nf-1)
                    am = re.match(r"^(\w+)\s*\((.*)\)$", item)
                    if am and am.group(1).lower() not in commons and am.group(1).lower() not in arrays:
                        dims = []
                        for d in split_top(am.group(2)):
                            dims.append("(%s, %s)" % tuple(pyexpr(x) for x in d.split(":")) if ":" in d else pyexpr(d))
                        emit("%s = FA(%s)" % (am.group(1).lower(), ", ".join(dims)))
                        known.add(am.group(1).lower())
            continue
        if re.match(r"^do\s+while\b", low):
            i = s.index("(")
            emit("while %s:" % pyexpr(s[i + 1:matching_paren(s, i)]))
            ind += 1
        elif re.match(r"^do\s+\w+\s*=", low):
            m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s.strip(), re.I)
            emit("for %s in frange(%s):" % (m.group(1).lower(), ", ".join(pyexpr(x) for x in split_top(m.group(2)))))
            ind += 1
        elif low in ("enddo", "end do", "endif", "end if"):
            ind -= 1
        elif low == "else":
            ind -= 1
            emit("else:")
            ind += 1
        elif low.startswith("if"):
            i = s.index("(")
            j = matching_paren(s, i)
            rest = s[j + 1:].strip()
            emit("if %s:" % pyexpr(s[i + 1:j]))
            ind += 1
            if rest.lower() != "then":
                stmt(rest)
                ind -= 1
        else:
            stmt(s)
    if name and result:
        emit("return _ret")
    return "\n".join(out) + "\n"


def namespace():
    ns = dict(FA=FA, frange=frange, min=min, max=max, abs=abs, exp=_intrinsic(math.exp), log=_intrinsic(math.log))
    from make_rconst_reference import f32
    ns["f32"] = f32
    # constants.f90:36-79
    ns.update(g=9.80665, gas_const=8.3144743, m_air=28.96546e-3, rhow=1000.0, avogadro=6.022140857e+23)
    ns["r0"] = ns["gas_const"] / ns["m_air"]
    return ns


def sources():
    src = {}
    src["vterm"] = translate(2793, 2864, "vterm", ("a", "t", "p"), result="vterm")
    src["advsed0"] = translate(5522, 5579, "advsed0", ("c", "y"), arrays=("c", "y"))
    src["advsed1"] = translate(5585, 5691, "advsed1", ("c", "y"), arrays=("c", "y"))
    cm = ("detw", "deta", "eta", "etw", "zb", "dzb", "dzbw", "tb", "eb", "ak", "d", "enw", "ew", "rn", "rw", "en", "e",
          "dew", "rq", "ff", "fsum", "nar", "theta", "thetl", "t", "talt", "p", "rho", "kw", "vt", "vd", "vdm", "rc",
          "sl1", "sion1")
    src["sedp"] = translate(2257, 2411, "sedp", ("dt",), arrays=cm, commons=("ajs", "trdep", "ds1", "ds2"))
    src["sedl"] = translate(2627, 2787, "sedl", ("dt",), arrays=cm)
    src["sedc_loop"] = translate(2567, 2596, arrays=("s1", "vg", "es1", "detw", "deta"))
    return src


def grid(nf, n, seed):
    r = np.random.default_rng(seed)
    detw = np.where(np.arange(n) < nf, 10.0, 10.0 * 1.1 ** (np.arange(n) - nf + 1.0))
    deta = 0.5 * (detw + np.roll(detw, -1)); deta[-1] = detw[-1]
    z = np.cumsum(detw)
    t = 288.0 - 0.0065 * z + r.uniform(-0.5, 0.5, n)
    p = 101325.0 * np.exp(-z / 8000.0)
    return detw, deta, t, p


def make(seed=7):
    nf, n, nka, nkt, nkc, j2, j6, j1 = 12, 15, 5, 6, 4, 4, 3, 6
    ncol = 3
    src = sources()
    ns = namespace()
    for k in ("vterm", "advsed0", "advsed1", "sedp", "sedl"):
        exec(src[k], ns)
    ns.update(nf=nf, n=n, nka=nka, nkt=nkt, nkc=nkc, j2=j2, j6=j6, nkc_l=3)
    r = np.random.default_rng(seed)
    out = dict(sizes=np.array([nf, n, nka, nkt, nkc, j2, j6, j1, 3]), dt=np.array(10.0))
    # advsed0 / advsed1 alone: profiles with zeros, steps and smooth parts; Courant numbers -1 < c <= 0
    ya = r.uniform(0.0, 1.0, (8, nf)) * 10.0 ** r.uniform(-3, 3, (8, 1))
    ya[1, 3:6] = 0.0; ya[2, :] = 1.0; ya[3, nf // 2:] = 0.0; ya[4] = 0.0
    ca = -r.uniform(0.0, 0.95, (8, nf)); ca[5] = 0.0; ca[6] = -0.999
    y0, y1 = ya.copy(), ya.copy()
    for i in range(8):
        ns["advsed0"](FA(nf, data=ca[i].copy()), FA(nf, data=y0[i]))
        ns["advsed1"](FA(nf, data=ca[i].copy()), FA(nf, data=y1[i]))
    out.update(adv_y=ya, adv_c=ca, adv_y0=y0, adv_y1=y1)
    # vterm over both regimes
    va = 10.0 ** r.uniform(-8, -3, 64); vt_t = r.uniform(230.0, 300.0, 64); vt_p = r.uniform(5.0e4, 1.02e5, 64)
    out.update(vterm_a=va, vterm_t=vt_t, vterm_p=vt_p, vterm=np.array([ns["vterm"](*x) for x in zip(va, vt_t, vt_p)]))
    # sedp: radii 0.01 um .. 500 um (both advection schemes, both vterm regimes, several sub-steps for the drops)
    rq = np.sort(10.0 ** r.uniform(-2, 2.7, (nka, nkt)), axis=1)
    e = 4.0 / 3.0 * np.pi * 1000.0 * (rq[0] * 1e-6) ** 3
    kw = r.integers(1, nkt + 1, nka).astype(np.int32)
    cols = dict(t=[], p=[], vd=[], ff0=[], ff1=[], diag0=[], diag1=[], rc=[], vt=[], vdm=[], sl0=[], sl1=[], si0=[], si1=[])
    for col in range(ncol):
        detw, deta, t, p = grid(nf, n, seed + col)
        ff = 10.0 ** r.uniform(-2, 3, (n, nka, nkt)) * (r.uniform(size=(n, nka, nkt)) < 0.8)
        ff[:, 1, 2] = 0.0                                   # an empty class: keeps the x0 of the class before it
        ff[:, 0, 0] = 1.0e-9                                # below the 1e-6 threshold, first class: x0 still 0
        vd = 10.0 ** r.uniform(-4, -1, (nka, nkt))
        diag = np.array([0.0, r.uniform(0, 1e-3), r.uniform(0, 1e-3), r.uniform(0, 1e-3)])
        ns.update(detw=FA(n, data=detw), deta=FA(n, data=deta), t=FA(n, data=t), p=FA(n, data=p),
                  rq=FA(nkt, nka, data=np.ascontiguousarray(rq.T)), e=FA(nkt, data=e), kw=FA(nka, data=kw),
                  vd=FA(nkt, nka, data=np.ascontiguousarray(vd.T)),
                  ajs=0.0, trdep=float(diag[1]), ds1=float(diag[2]), ds2=float(diag[3]))
        fft = np.ascontiguousarray(ff.transpose(2, 1, 0))   # ff(nkt,nka,n)
        ns["ff"] = FA(nkt, nka, n, data=fft)
        cols["ff0"].append(ff.copy())
        ns["x0"] = 0.0
        # x0 is an uninitialised local in the reference; every evaluation here starts it at 0 (include/mistra_sed.h)
        code = src["sedp"].replace("    ajs =  0.0\n", "    ajs =  0.0\n    x0 = 0.0\n", 1)
        assert code != src["sedp"]
        exec(code, ns)
        ns["sedp"](10.0)
        cols["ff1"].append(fft.transpose(2, 1, 0).copy())
        cols["diag0"].append(diag)
        cols["diag1"].append(np.array([ns["ajs"], ns["trdep"], ns["ds1"], ns["ds2"]]))
        cols["t"].append(t); cols["p"].append(p); cols["vd"].append(vd)
        # sedl
        rc = 10.0 ** r.uniform(-8.5, -3.5, (n, nkc))
        vt = 10.0 ** r.uniform(-6, 0, (nf, nkc))
        vdm = 10.0 ** r.uniform(-4, -1, nkc)
        sl = 10.0 ** r.uniform(-12, -6, (n, nkc, j2)) * (r.uniform(size=(n, nkc, j2)) < 0.85)
        si = 10.0 ** r.uniform(-12, -6, (n, nkc, j6)) * (r.uniform(size=(n, nkc, j6)) < 0.85)
        slt, sit = np.ascontiguousarray(sl.transpose(2, 1, 0)), np.ascontiguousarray(si.transpose(2, 1, 0))
        ns.update(rc=FA(nkc, n, data=np.ascontiguousarray(rc.T)), vt=FA(nkc, nf, data=np.ascontiguousarray(vt.T)),
                  vdm=FA(nkc, data=vdm), sl1=FA(j2, nkc, n, data=slt), sion1=FA(j6, nkc, n, data=sit))
        ns["sedl"](10.0)
        vtn = np.zeros((n, nkc)); vtn[:nf] = vt
        cols["rc"].append(rc); cols["vt"].append(vtn); cols["vdm"].append(vdm)
        cols["sl0"].append(sl); cols["sl1"].append(slt.transpose(2, 1, 0).copy())
        cols["si0"].append(si); cols["si1"].append(sit.transpose(2, 1, 0).copy())
    out.update({k: np.array(v) for k, v in cols.items()})
    out.update(detw=detw, deta=deta, rq=rq, e=e, kw=kw)
    # sedc species loop
    vg = np.array([0.0, 0.5e-5, 1.0e-5, 0.27e-2, 0.1e-1, 0.2e-2])
    es1 = np.array([0.0, 1.0e9, 0.0, 3.0e10, 0.0, 1.0e12])
    s1 = 10.0 ** r.uniform(-12, -7, (ncol, n, j1))
    s1n = s1.copy()
    for col in range(ncol):
        st = np.ascontiguousarray(s1n[col].T)               # s1(j1,n)
        ns.update(j1=j1, vg=FA(j1, data=vg), es1=FA(j1, data=es1), s1=FA(j1, n, data=st), dt=10.0)
        exec(src["sedc_loop"], ns)
        s1n[col] = st.T
    out.update(sedc_vg=vg, sedc_es1=es1, sedc_s0=s1, sedc_s1=s1n)
    return out, src


if __name__ == "__main__":
    fx, src = make()
    if "--show" in sys.argv:
        for k, v in src.items():
            print("# ---- %s\n%s" % (k, v))
    path = os.path.join(HERE, "sed_reference.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path), "bytes")
