"""Second opinion on the CPU oracle's generated blocks: Fun_x, Jac_SP_x and KppSolve_x evaluated by EXECUTING THE
REFERENCE'S OWN Fortran statements (authoring container only, needs /root/reference):

    python tests/golden/make_kpp_blocks_reference.py     # writes tests/golden/kpp_blocks_reference_{gas,aer,tot}.npz

The statements of SUBROUTINE Fun_x / Jac_SP_x / KppSolve_x (gas.f:2043-2498, 2656-6100, 6206-6636 and the aer.f / tot.f
counterparts) are plain assignments  NAME(i) = expression  over the arrays A, B, V, F, RCT, Vdot, JVS, X.  They are read
from the Fortran text (continuation lines joined), array references turned into Python indexing, literals typed as in
tests/golden/make_rconst_reference.py (default-REAL coefficients such as 0.05*A(135) are binary32 values promoted when
they meet a double), and executed with Python floats = IEEE binary64 without fused multiply-add, in the reference's
statement order.  No code is shared with mistra_b200/mechgen/extract.py or oracle/emit_oracle.py.

tests/test_kpp_blocks_reference.py requires the oracle's fun / jac / solve to reproduce the stored results BIT FOR BIT;
tests/test_codegen_host.py and tests/test_onchip_tables.py tie the CUDA code generators to the oracle in the same way.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_rconst_reference import REF, expr, f32, parameters, strip_comment  # noqa: E402

ARRAYS = ("a", "b", "v", "f", "rct", "vdot", "jvs", "x")


def subroutine_statements(mech, name):
    """Assignment statements (python source lines) of SUBROUTINE <name> in <mech>.f."""
    lines = open(os.path.join(REF, "%s.f" % mech), errors="replace").read().split("\n")
    a = next(i for i, l in enumerate(lines) if re.match(r"^\s+SUBROUTINE %s\b" % name, l))
    b = next(i for i in range(a, len(lines)) if re.match(r"^\s+END\s*$", lines[i]))
    logical = []
    for ln in lines[a + 1:b]:
        if not ln.strip() or ln[0] in "Cc!*":
            continue
        ln = strip_comment(ln)
        if len(ln) > 5 and ln[5] not in " 0" and ln[:5].strip() == "":
            logical[-1] += ln[6:].strip()
        else:
            logical.append(ln.strip())
    out = []
    for s in logical:
        m = re.match(r"^(\w+)\((\d+)\)\s*=\s*(.*)$", s)
        if not m or m.group(1).lower() not in ARRAYS:
            continue                                      # declarations, INCLUDE, RETURN
        rhs = expr(m.group(3))
        rhs = re.sub(r"\b(%s)\((\d+)\)" % "|".join(ARRAYS), lambda q: "%s[%s]" % (q.group(1), q.group(2)), rhs)
        out.append("%s[%s] = %s" % (m.group(1).lower(), m.group(2), rhs))
    return out


def run(stmts, **arrays):
    ns = {"f32": f32}
    ns.update(arrays)
    exec(compile("\n".join(stmts), "<reference statements>", "exec"), ns)
    return ns


def one_based(a):
    return [0.0] + [float(x) for x in a]


def main():
    from tests import util
    for mech in ("gas", "aer", "tot"):
        x = mech[0]
        par = parameters(mech)
        nvar, nfix, nreact, nnz = par["nvar"], par["nfix"], par["nreact"], par["lu_nonzero"]
        fun = subroutine_statements(mech, "Fun_%s" % x)
        jac = subroutine_statements(mech, "Jac_SP_%s" % x)
        sol = subroutine_statements(mech, "KppSolve_%s" % x)
        ncell = 3
        var, fix, rc = util.random_cells(mech, ncell, 9100 + len(mech))
        r = np.random.default_rng(77)
        vdot = np.zeros((ncell, nvar))
        jvs = np.zeros((ncell, nnz))
        lu = r.normal(size=(ncell, nnz))
        lu[np.abs(lu) < 0.2] += 1.0                    # KppSolve divides by the diagonal entries
        xin = r.normal(size=(ncell, nvar))
        xout = np.zeros((ncell, nvar))
        for c in range(ncell):
            ns = run(fun, v=one_based(var[c]), f=one_based(fix[c]), rct=one_based(rc[c]), a=[0.0] * (nreact + 1),
                     vdot=[0.0] * (nvar + 1))
            vdot[c] = ns["vdot"][1:]
            nb = max(int(q) for s in jac for q in re.findall(r"\bb\[(\d+)\]", s))
            ns = run(jac, v=one_based(var[c]), f=one_based(fix[c]), rct=one_based(rc[c]), b=[0.0] * (nb + 1),
                     jvs=[0.0] * (nnz + 1))
            jvs[c] = ns["jvs"][1:]
            ns = run(sol, jvs=one_based(lu[c]), x=one_based(xin[c]))
            xout[c] = ns["x"][1:]
        assert np.isfinite(vdot).all() and np.isfinite(jvs).all() and np.isfinite(xout).all()
        out = os.path.join(HERE, "kpp_blocks_reference_%s.npz" % mech)
        np.savez_compressed(out, var=var, fix=fix, rconst=rc, vdot=vdot, jvs=jvs, lu=lu, xin=xin, xout=xout)
        print("%s: Fun %d, Jac_SP %d, KppSolve %d statements executed for %d cells -> %s"
              % (mech, len(fun), len(jac), len(sol), ncell, out))


if __name__ == "__main__":
    main()
