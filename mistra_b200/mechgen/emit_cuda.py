"""Emit the CUDA device code of one mechanism from its tables.

Output: ``mistra_b200/csrc/_gen/mech_<x>.cuh`` (x = g, a, t) with straight-line
device functions over a per-lane workspace ``w`` (one cell per thread, SoA
interleaved by lane: element ``e`` of this lane's cell sits at ``w[e*32]``, so a
warp access is one aligned 256-byte segment and every address is
``lane_base + immediate``):

* ``fun_<x><MODE>``     reaction rates + aggregation (role of Fun_x, gas.f:2043);
* ``jacprep_<x>``       sparse Jacobian fused with ros_PrepareMatrix's
                        ``Ghimj = -Jac0; diag += 1/(H*gamma)`` (Jac_SP_x gas.f:2656
                        + gas.f:1444-1449), returns the singularity flag of
                        KppDecomp's diagonal test (gas.f:6156);
* ``decomp_<x>``        code-generated row-wise sparse LU without pivoting on the
                        fixed pattern (role of KppDecomp_x, gas.f:6142), storing
                        reciprocal pivots in the diagonal slots;
* ``solve_<x><MODE>``   forward/backward substitution (role of KppSolve_x,
                        gas.f:6206) with the stage-combination epilogues of
                        RosenbrockIntegrator_x (gas.f:1247-1294) fused into the
                        backward sweep.

Operation order inside every sum and every elimination follows the reference
statement order, so the only rounding differences against the CPU oracle are
FMA contraction, reciprocal-pivot multiplication and the fused error-norm order
(all removable with -DKPP_STRICT for the bit-parity test build).
"""
from __future__ import annotations

import os
import sys

from . import mech as mechmod


def greedy_order(uses):
    """uses[s] = set of temporaries statement s reads.  Returns an order of the
    statements that keeps few temporaries live at once (each temporary is
    computed right before its first use and dies after its last): repeatedly
    pick the statement with the smallest (new temporaries - temporaries it
    releases)."""
    n = len(uses)
    remaining = {}
    for s, u in enumerate(uses):
        for t in u:
            remaining[t] = remaining.get(t, 0) + 1
    live = set()
    todo = set(range(n))
    order = []
    # statements without temporaries first (they cost nothing)
    for s in sorted(todo):
        if not uses[s]:
            order.append(s)
    todo -= set(order)
    users = {}
    for s in todo:
        for t in uses[s]:
            users.setdefault(t, set()).add(s)
    while todo:
        # candidates: statements touching a live temporary, else all
        cand = set()
        for t in live:
            cand |= users[t] & todo
        if not cand:
            cand = todo
        best, bkey = None, None
        for s in cand:
            new = sum(1 for t in uses[s] if t not in live)
            rel = sum(1 for t in uses[s] if remaining[t] == 1)
            key = (new - rel, new, s)
            if bkey is None or key < bkey:
                best, bkey = s, key
        order.append(best)
        todo.discard(best)
        for t in uses[best]:
            live.add(t)
            remaining[t] -= 1
            if remaining[t] == 0:
                live.discard(t)
    return order


class Emitter:
    def __init__(self, m, unroll_decomp, tail=0):
        self.m = m
        self.tail = tail
        # structural fill-in of the LU pattern: Jac_SP sets these to 0 (gas.f:2675-6100), so after
        # Ghimj = -Jac0 they are -0.0 - no need to store and re-load them
        diag = set(int(d) for d in m.diag[:m.nvar])
        self.fill = set(nz for nz in range(m.lu_nonzero) if not m.jvs[nz] and nz not in diag)
        touched = set(op[2] for op in m.decomp_ops() if op[0] == "upd")
        assert self.fill <= touched, "a fill-in entry that no elimination step writes"
        self.x = m.suffix
        self.unroll_decomp = unroll_decomp
        self.coef = {c: i for i, c in enumerate(c for c in m.coef_literals if not c.isdigit())}
        self.lines = []

    def w(self, s=""):
        self.lines.append(s)

    # -- expression helpers ---------------------------------------------------
    def lit(self, c):
        if c.isdigit():
            return "%s.0" % c
        return "CF(%d)" % self.coef[c]

    def prod(self, facs, vmac):
        out = []
        for kind, v in facs:
            if kind == "N":
                out.append(self.lit(v))
            elif kind == "R":
                out.append("R_(%d)" % v)
            elif kind == "V":
                out.append("%s(%d)" % (vmac, v))
            else:
                out.append("F_(%d)" % v)
        return "*".join(out)

    def summ(self, terms, sym):
        if not terms:
            return None
        s = ""
        for i, (sign, coef, idx) in enumerate(terms):
            t = "%s%d" % (sym, idx)
            if coef is not None:
                t = "%s*%s" % (self.lit(coef), t)
            if i == 0:
                s += ("-" if sign < 0 else "") + t
            else:
                s += (" - " if sign < 0 else " + ") + t
        return s

    # -- pieces ---------------------------------------------------------------
    def emit_header(self):
        m, x = self.m, self.x
        w = self.w
        w("// GENERATED by mistra_b200/mechgen/emit_cuda.py from mistra_b200/mech/%s.json - do not edit." % m.name)
        w("// Mechanism '%s': NVAR=%d NFIX=%d NREACT=%d LU_NONZERO=%d" % (
            m.name, m.nvar, m.nfix, m.nreact, m.lu_nonzero))
        w("#pragma once")
        w("namespace mech_%s {" % x)
        w("constexpr int NVAR = %d, NFIX = %d, NREACT = %d, LU_NONZERO = %d, NCOEF = %d;" % (
            m.nvar, m.nfix, m.nreact, m.lu_nonzero, max(1, len(self.coef))))
        # slots (units: doubles per lane)
        w("constexpr int S_Y = 0, S_K1 = NVAR, S_K2 = 2 * NVAR, S_K3 = 3 * NVAR, S_T1 = 4 * NVAR;")
        w("constexpr int S_FIX = 5 * NVAR, S_RCT = S_FIX + NFIX, S_G = S_RCT + NREACT;")
        w("constexpr int NSLOT = S_G + LU_NONZERO;")
        w("__constant__ double c_coef[NCOEF];")
        lits = [c for c in self.m.coef_literals if not c.isdigit()]
        w("static const char *const coef_literals[NCOEF] = {%s};" % (
            ", ".join('"%s"' % c for c in lits) if lits else '"0"'))
        w("#define CF(i) c_coef[i]")
        w("#define EL(s, i) w[((s) + (i)) * 32]")
        w("#define R_(i) EL(S_RCT, i)")
        w("#define F_(i) EL(S_FIX, i)")
        w("#define G_(i) EL(S_G, i)")
        w("#define ROW_FENCE() asm volatile(\"\" ::: \"memory\")")
        w("#ifndef KPP_STRICT   // diagonal slots hold reciprocal pivots")
        w("#define PIV_SCALE(r, d) r *= (d)")
        w("#define PIV_FINISH(i, v) G_(i) = 1.0 / (v)")
        w("#define PIV_APPLY(s, d) (s) * (d)")
        w("#else                // divisions exactly as KppDecomp/KppSolve")
        w("#define PIV_SCALE(r, d) r /= (d)")
        w("#define PIV_FINISH(i, v) G_(i) = (v)")
        w("#define PIV_APPLY(s, d) (s) / (d)")
        w("#endif")
        w("// Stage-combination epilogues fused into the backward sweep of KppSolve")
        w("// (gas.f:1254-1258, 1266-1268, 1281-1294; Ros3 M/E coefficients gas.f:1611-1617).")
        w("#ifdef KPP_STRICT")
        w("#define KPP_PARK_YERR(k, ye) EL(S_K2, k) = (ye);")
        w("#else")
        w("#define KPP_PARK_YERR(k, ye)")
        w("#endif")
        w("#define SOLVE_EPILOGUE(k) \\")
        w("  X_(k) = x; \\")
        w("  if (MODE == 1) { EL(S_T1, k) = EL(S_Y, k) + x; } \\")
        w("  else if (MODE == 2) { EL(S_K3, k) = EL(S_K3, k) + hc32 * x; } \\")
        w("  else { \\")
        w("    const double y = EL(S_Y, k), k1 = EL(S_K1, k), k2 = EL(S_K2, k); \\")
        w("    const double yn = ((y + k1) + 0.61697947043828245592553615689730e+01 * k2) \\")
        w("                      + (-0.42772256543218573326238373806514e+00) * x; \\")
        w("    const double ye = (0.5 * k1 + (-0.29079558716805469821718236208017e+01) * k2) \\")
        w("                      + 0.22354069897811569627360909276199e+00 * x; \\")
        w("    const double q = ye / (atol + rtol * fmax(fabs(y), fabs(yn))); \\")
        w("    esum = esum + q * q; \\")
        w("    EL(S_T1, k) = yn; \\")
        w("    KPP_PARK_YERR(k, ye) \\")
        w("  }")
        w("")

    def emit_fun(self):
        m, x = self.m, self.x
        w = self.w
        # MODE 0: V = Y,  K1 = f
        # MODE 1: V = T1, K2 = f + hc21*K1 ; K3 = f + hc31*K1   (gas.f:1263-1268)
        w("template <int MODE>")
        w("__device__ __noinline__ void fun(double *__restrict__ w, double hc21, double hc31)")
        w("{")
        w("#define V_(i) EL(MODE == 0 ? S_Y : S_T1, i)")
        uses = [set(idx for _, _, idx in terms) for terms in m.vdot]
        done = set()
        for i in greedy_order(uses):
            terms = m.vdot[i]
            for _, _, idx in terms:
                if idx not in done:
                    done.add(idx)
                    w("  const double a%d = %s;" % (idx, self.prod(m.reactions[idx], "V_")))
            e = self.summ(terms, "a") or "0.0"
            w("  { const double f = %s;" % e)
            w("    if (MODE == 0) { EL(S_K1, %d) = f; }" % i)
            w("    else { const double k1 = EL(S_K1, %d); EL(S_K2, %d) = f + hc21 * k1; EL(S_K3, %d) = f + hc31 * k1; } }" % (i, i, i))
        w("#undef V_")
        w("}")
        w("")

    def emit_jacprep(self):
        m = self.m
        w = self.w
        w("// G = -Jac_SP(Y); G[diag] += ghinv.  Returns nonzero if a prepared diagonal is exactly 0.")
        w("__device__ __noinline__ int jacprep(double *__restrict__ w, double ghinv)")
        w("{")
        w("#define V_(i) EL(S_Y, i)")
        w("  int sing = 0;")
        Bd = {k: facs for k, facs in m.B}
        diagset = {int(m.diag[k]): k for k in range(m.nvar)}
        uses = [set(idx for _, _, idx in terms) for terms in m.jvs]
        done = set()
        for nz in greedy_order(uses):
            terms = m.jvs[nz]
            for _, _, idx in terms:
                if idx not in done:
                    done.add(idx)
                    w("  const double b%d = %s;" % (idx, self.prod(Bd[idx], "V_")))
            e = self.summ(terms, "b")
            if nz in diagset:
                if e is None:
                    w("  { const double g = -0.0 + ghinv; G_(%d) = g; sing |= (g == 0.0); }" % nz)
                else:
                    w("  { const double g = -(%s) + ghinv; G_(%d) = g; sing |= (g == 0.0); }" % (e, nz))
            else:
                if e is None:
                    assert nz in self.fill
                    if self.unroll_decomp != "tiled":   # the tiled LU creates fill-in in registers (gload)
                        w("  G_(%d) = -0.0;" % nz)
                else:
                    w("  G_(%d) = -(%s);" % (nz, e))
        w("#undef V_")
        w("  return sing;")
        w("}")
        w("")

    def emit_decomp_unrolled(self):
        m = self.m
        w = self.w
        w("// Row-wise sparse LU on the fixed pattern; diagonal slots end up holding 1/pivot")
        w("// (KPP_STRICT: the pivot itself, divisions as in KppDecomp).")
        w("__device__ __noinline__ void decomp(double *__restrict__ w)")
        w("{")
        for k in range(m.nvar):
            lo, dg, hi = int(m.crow[k]), int(m.diag[k]), int(m.crow[k + 1])
            if dg == lo:
                w("  PIV_FINISH(%d, G_(%d));" % (dg, dg))
                continue
            w("  {")
            pos = {int(m.icol[kk]): kk for kk in range(lo, hi)}
            # program of the row: list of (kind, args) in reference order
            prog = []
            for kk in range(lo, dg):
                j = int(m.icol[kk])
                dj = int(m.diag[j])
                prog.append(("piv", kk, dj))
                for jj in range(dj + 1, int(m.crow[j + 1])):
                    prog.append(("upd", pos[int(m.icol[jj])], kk, jj))
            last = {}
            for n, op in enumerate(prog):
                for r in ((op[1],) if op[0] == "piv" else (op[1], op[2])):
                    last[r] = n
            last[dg] = max(last.get(dg, -1), len(prog) - 1) if dg in last else -1
            loaded = set()

            def need(r):
                if r not in loaded:
                    loaded.add(r)
                    w("    double r%d = G_(%d);" % (r, r))

            def retire(n):
                for r in [r for r in loaded if last.get(r, -1) == n and r != dg]:
                    w("    G_(%d) = r%d;" % (r, r))
            for n, op in enumerate(prog):
                if op[0] == "piv":
                    need(op[1])
                    w("    PIV_SCALE(r%d, G_(%d));" % (op[1], op[2]))
                else:
                    need(op[1])
                    w("    r%d -= r%d * G_(%d);" % (op[1], op[2], op[3]))
                retire(n)
            if dg in loaded:
                w("    PIV_FINISH(%d, r%d);" % (dg, dg))
            else:
                w("    PIV_FINISH(%d, G_(%d));" % (dg, dg))
            # keep the compiler from caching U-row loads in registers across rows
            # (they would only be spilled to local memory again)
            w("    ROW_FENCE();")
            w("  }")
        w("}")
        w("")

    def gload(self, idx):
        """First read of an LU slot by the factorisation: fill-in starts as -0.0."""
        return "-0.0" if idx in self.fill else "G_(%d)" % idx

    # -- tile-blocked LU ---------------------------------------------------------
    def _emit_row_program(self, k, head_only_below=None):
        """Row-wise elimination of row k (KppDecomp order).  With head_only_below = h
        only pivots j < h and destinations c < h are processed (the tail rows' part
        that lies in the head columns); the rest of that row is done by the tiles."""
        m, w = self.m, self.w
        lo, dg, hi = int(m.crow[k]), int(m.diag[k]), int(m.crow[k + 1])
        h = head_only_below
        if h is None and dg == lo:
            w("  PIV_FINISH(%d, G_(%d));" % (dg, dg))
            return
        pos = {int(m.icol[kk]): kk for kk in range(lo, hi)}
        prog = []
        for kk in range(lo, dg):
            j = int(m.icol[kk])
            if h is not None and j >= h:
                break
            dj = int(m.diag[j])
            prog.append(("piv", kk, dj))
            for jj in range(dj + 1, int(m.crow[j + 1])):
                c = int(m.icol[jj])
                if h is not None and c >= h:
                    continue
                prog.append(("upd", pos[c], kk, jj))
        if not prog:
            return
        last = {}
        for n, op in enumerate(prog):
            for r in ((op[1],) if op[0] == "piv" else (op[1], op[2])):
                last[r] = n
        if h is None:
            last[dg] = -1
        loaded = set()
        w("  {")

        def need(r):
            if r not in loaded:
                loaded.add(r)
                w("    double r%d = %s;" % (r, self.gload(r)))

        for n, op in enumerate(prog):
            if op[0] == "piv":
                need(op[1])
                w("    PIV_SCALE(r%d, G_(%d));" % (op[1], op[2]))
            else:
                need(op[1])
                w("    r%d -= r%d * G_(%d);" % (op[1], op[2], op[3]))
            for r in [r for r in loaded if last.get(r, -1) == n]:
                w("    G_(%d) = r%d;" % (r, r))
        if h is None:
            if dg in loaded:
                w("    PIV_FINISH(%d, r%d);" % (dg, dg))
            else:
                w("    PIV_FINISH(%d, G_(%d));" % (dg, dg))
        w("    ROW_FENCE();")
        w("  }")

    def emit_decomp_tiled(self, tail, B=8):
        """Blocked (Crout-order) LU of the trailing `tail` x `tail` region in B x B
        register tiles; head rows row-wise as in emit_decomp_unrolled.  Every entry
        (i,j) still receives its updates L(i,k)*U(k,j) one by one in increasing k
        (the order of KppDecomp, gas.f:6142-6177), so the factors are the same to the
        last bit; what changes is that a tile's entries stay in registers over the
        whole k range and each loaded L/U value feeds up to B multiply-adds."""
        m, w = self.m, self.w
        n = m.nvar
        tail = min(tail, n) // B * B
        h = n - tail
        nt = tail // B
        pos = m.pos
        blk = lambda r: (r - h) // B
        rows_of_col = {}
        cols_of_row = {}
        for (i, c) in pos:
            if c < i:
                rows_of_col.setdefault(c, []).append(i)
            elif c > i:
                cols_of_row.setdefault(i, []).append(c)
        funcs = []
        # pass 0: head rows, then the head-column part of the tail rows
        w("// tile-blocked LU: head = rows [0,%d) row-wise, tail = %d x %d in %dx%d tiles" % (h, tail, tail, B, B))
        w("__device__ __noinline__ void decomp_p0(double *__restrict__ w)")
        w("{")
        for k in range(h):
            self._emit_row_program(k)
        for k in range(h, n):
            self._emit_row_program(k, head_only_below=h)
        w("}")
        w("")
        funcs.append("decomp_p0")
        for I in range(nt):
            fname = "decomp_t%d" % I
            funcs.append(fname)
            w("__device__ __noinline__ void %s(double *__restrict__ w)" % fname)
            w("{")
            R = list(range(h + I * B, h + I * B + B))
            for J in range(nt):
                Cc = list(range(h + J * B, h + J * B + B))
                ent = {(i, j): pos[(i, j)] for i in R for j in Cc if (i, j) in pos}
                if not ent:
                    continue
                reg = {e: "c%d" % idx for e, idx in ent.items()}
                w("  {  // tile (%d,%d): %d entries" % (I, J, len(ent)))
                for e, idx in sorted(ent.items(), key=lambda t: t[1]):
                    w("    double %s = %s;" % (reg[e], self.gload(idx)))
                # candidate pivots
                ks = set()
                for i in R:
                    for c in range(int(m.crow[i]), int(m.diag[i])):
                        ks.add(int(m.icol[c]))
                if I == J:
                    ks |= set(R[:-1])           # every pivot of a diagonal tile is finished here
                for k in sorted(ks):
                    if k >= R[-1] or k > Cc[-1]:
                        break
                    K = blk(k) if k >= h else -1
                    rows_k = [i for i in R if i > k and (i, k) in pos]
                    cols_k = [j for j in Cc if j > k and (k, j) in pos]
                    in_tile_col = (K == J)      # L(i,k) is an entry of this tile
                    in_tile_row = (K == I)      # U(k,j) is an entry of this tile
                    if in_tile_col and in_tile_row:
                        d = int(m.diag[k])
                        w("    PIV_FINISH(%d, %s);" % (d, reg[(k, k)]))
                    if not rows_k:
                        continue
                    if in_tile_col and in_tile_row:
                        w("    { const double pv = G_(%d);" % d)
                        for i in rows_k:
                            w("      PIV_SCALE(%s, pv);" % reg[(i, k)])
                        w("    }")
                    elif in_tile_col:
                        d = int(m.diag[k])
                        w("    { const double pv = G_(%d);" % d)
                        for i in rows_k:
                            w("      PIV_SCALE(%s, pv);" % reg[(i, k)])
                        w("    }")
                    if not cols_k:
                        continue
                    w("    {")
                    ln, un = {}, {}
                    for i in rows_k:
                        if in_tile_col:
                            ln[i] = reg[(i, k)]
                        else:
                            ln[i] = "l%d" % i
                            w("      const double l%d = G_(%d);" % (i, pos[(i, k)]))
                    for j in cols_k:
                        if in_tile_row:
                            un[j] = reg[(k, j)]
                        else:
                            un[j] = "u%d" % j
                            w("      const double u%d = G_(%d);" % (j, pos[(k, j)]))
                    for i in rows_k:
                        for j in cols_k:
                            assert (i, j) in ent, (i, j, k)
                            w("      %s -= %s * %s;" % (reg[(i, j)], ln[i], un[j]))
                    w("    }")
                # remaining pivots of a diagonal tile's last row / store
                if I == J:
                    k = R[-1]
                    w("    PIV_FINISH(%d, %s);" % (int(m.diag[k]), reg[(k, k)]))
                for e, idx in sorted(ent.items(), key=lambda t: t[1]):
                    if I == J and e[0] == e[1]:
                        continue
                    w("    G_(%d) = %s;" % (idx, reg[e]))
                w("    ROW_FENCE();")
                w("  }")
            w("}")
            w("")
        w("__device__ __forceinline__ void decomp(double *__restrict__ w)")
        w("{")
        for f in funcs:
            w("  %s(w);" % f)
        w("}")
        w("")

    def emit_decomp_tables(self):
        """Index tables for the loop-driven decomposition (warp-uniform indices,
        read through the constant cache)."""
        m = self.m
        w = self.w

        def arr(ty, name, a):
            w("__constant__ %s %s[%d] = {" % (ty, name, len(a)))
            for i in range(0, len(a), 24):
                w("  " + ",".join(str(int(v)) for v in a[i:i + 24]) + ",")
            w("};")
        arr("unsigned short", "c_icol", m.icol)
        arr("unsigned short", "c_crow", m.crow)
        arr("unsigned short", "c_diag", m.diag)
        w("")

    def emit_solve(self):
        m = self.m
        w = self.w
        # MODE 1: X = K1 ; epilogue T1 = Y + X                      (gas.f:1254-1258, A21 = 1)
        # MODE 2: X = K2 ; epilogue K3 += hc32 * X                  (gas.f:1266-1268)
        # MODE 3: X = K3 ; epilogue Ynew, Yerr, error-norm sum      (gas.f:1281-1294)
        w("template <int MODE>")
        w("__device__ __noinline__ double solve(double *__restrict__ w, double hc32, double atol, double rtol)")
        w("{")
        w("#define X_(i) EL(MODE == 1 ? S_K1 : (MODE == 2 ? S_K2 : S_K3), i)")
        w("  double esum = 0.0;")

        def XR(i):
            return "X_(%d)" % int(i)

        def fwd_row(k):
            lo, dg = int(m.crow[k]), int(m.diag[k])
            if dg > lo:
                s = "  %s = %s" % (XR(k), XR(k))
                for kk in range(lo, dg):
                    s += " - G_(%d) * %s" % (kk, XR(m.icol[kk]))
                w(s + ";")

        def bwd_row(k):
            dg, hi = int(m.diag[k]), int(m.crow[k + 1])
            s = XR(k)
            for kk in range(dg + 1, hi):
                s += " - G_(%d) * %s" % (kk, XR(m.icol[kk]))
            w("  {")
            w("    const double x = PIV_APPLY(%s, G_(%d));" % (s, dg))
            w("    SOLVE_EPILOGUE(%d)" % k)
            w("  }")

        for k in range(m.nvar):
            fwd_row(k)
        for k in range(m.nvar - 1, -1, -1):
            bwd_row(k)
        w("#undef X_")
        w("  return esum;")
        w("}")
        w("")

    def emit(self):
        self.emit_header()
        self.emit_fun()
        self.emit_jacprep()
        if self.unroll_decomp == "tiled":
            self.w("#define MECH_DECOMP_UNROLLED 1")
            self.emit_decomp_tiled(self.tail)
        elif self.unroll_decomp:
            self.w("#define MECH_DECOMP_UNROLLED 1")
            self.emit_decomp_unrolled()
        self.emit_decomp_tables()
        self.emit_solve()
        self.w("}  // namespace mech_%s" % self.x)
        self.w("#undef CF\n#undef R_\n#undef F_")
        return "\n".join(self.lines) + "\n"


def main(argv):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outdir = os.path.join(root, "mistra_b200", "csrc", "_gen")
    for a in argv[1:]:
        if a.startswith("--outdir="):
            outdir = a.split("=", 1)[1]
    os.makedirs(outdir, exist_ok=True)
    # LU variant per mechanism: "tiled" = row-wise head + register-tiled trailing block
    # of the given size (covers >= 96 % of the multiply-adds)
    unroll = {"gas": "tiled", "aer": "tiled", "tot": "tiled"}
    tail = {"gas": 32, "aer": 96, "tot": 128}
    for a in argv[1:]:
        if a.startswith("--lu="):           # e.g. --lu=gas:rows,aer:loop
            for kv in a.split("=", 1)[1].split(","):
                n, v = kv.split(":")
                unroll[n] = {"rows": True, "loop": False, "tiled": "tiled"}[v]
        if a.startswith("--tail="):
            for kv in a.split("=", 1)[1].split(","):
                n, v = kv.split(":")
                tail[n] = int(v)
    for name in mechmod.MECH_NAMES:
        m = mechmod.load(name)
        text = Emitter(m, unroll[name], tail[name]).emit()
        path = os.path.join(outdir, "mech_%s.cuh" % m.suffix)
        if not (os.path.exists(path) and open(path).read() == text):
            with open(path, "w") as f:
                f.write(text)
        print("wrote", path, len(text) // 1024, "KiB")
    # Update_RCONST_x right-hand sides (gas.f:316-664, aer.f:345-1398, tot.f:1081-2803)
    for name in mechmod.MECH_NAMES:
        m = mechmod.load(name)
        text = "// GENERATED by mistra_b200/mechgen/emit_cuda.py - RCONST(i) of Update_RCONST_%s (%s)\n" % (
            m.suffix, m.source["Update_RCONST"])
        text += "".join("RC[%d] = %s;\n" % (i, e) for i, e in enumerate(m.d["rconst"]))
        path = os.path.join(outdir, "rconst_%s.inc" % m.suffix)
        if not (os.path.exists(path) and open(path).read() == text):
            with open(path, "w") as f:
                f.write(text)
    # species-name table for mistra_kpp_spc_name (SPC_NAMES, gas.f:6867-6891)
    lines = ["// GENERATED by mistra_b200/mechgen/emit_cuda.py - do not edit.", "#include <cstddef>"]
    for name in mechmod.MECH_NAMES:
        m = mechmod.load(name)
        lines.append("static const char *const names_%s[%d] = {" % (m.suffix, m.nspec))
        for i in range(0, m.nspec, 8):
            lines.append("  " + ", ".join('"%s"' % n for n in m.spc_names[i:i + 8]) + ",")
        lines.append("};")
    lines.append("extern \"C\" const char *mistra_kpp_spc_name_impl(int mech, int i)\n{")
    for name in mechmod.MECH_NAMES:
        m = mechmod.load(name)
        lines.append("  if (mech == %d) return (i >= 0 && i < %d) ? names_%s[i] : NULL;" % (
            mechmod.MECH_ID[name], m.nspec, m.suffix))
    lines.append("  return NULL;\n}")
    text = "\n".join(lines) + "\n"
    path = os.path.join(outdir, "kpp_names.cpp")
    if not (os.path.exists(path) and open(path).read() == text):
        with open(path, "w") as f:
            f.write(text)
    print("wrote", path)


if __name__ == "__main__":
    main(sys.argv)
