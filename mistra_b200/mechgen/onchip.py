"""Tables and code of the ON-CHIP Ros3 kernels (one thread block per cell).

Design: a cell's whole integrator state stays on the SM for the
10 s call.  The LU factors of ``Ghimj`` (role of KppDecomp_x / KppSolve_x,
/root/reference/src/gas.f:6142-6636, aer.f:20959-, tot.f:39468-) are split at row
``h = NVAR - T``:

* the *head* (rows/columns < h, plus the head-column part of the last T rows and the
  tail-column part of the head rows) is sparse and irregular; it lives in shared
  memory as a compact list ``G`` (head diagonals first, then the reference's CSR order
  without the tail block) and is processed by small table-driven interpreters whose
  instruction streams come from global memory through the block's TMA pipeline (one copy serves all cell
  slots of the block);
* the *tail* (the trailing T x T block, 83-95 % dense after fill-in) lives in
  REGISTERS: lane p of warp w holds rows ``h+p+32q`` (q < R) and columns
  ``h+32w .. h+32w+31`` as ``a[q][s]``; its elimination, the head->tail updates and
  the triangular sweeps are unrolled code emitted here.

Every LU entry receives its updates ``-= L(i,j)*U(j,c)`` in increasing j, i.e. in the
order of KppDecomp's row-wise loop, so the factors are the reference's to the last
bit when compiled without FMA contraction and with true divisions (KPP_STRICT).

This module builds the tables (numpy), emulates every interpreter on the CPU in
exactly the order the kernel uses (tests/test_onchip_tables.py compares that with the
CPU oracle bit for bit) and emits ``csrc/_gen/onchip_<x>.cuh``.

Stream formats (all indices are double-word indices into the block's shared memory,
pre-shifted by 3 so that the low bits carry flags):

* product ops (Fun / Jac_SP rate products), two 32-bit words:
  ``w0 = r | f1<<11 | f2<<20``, ``w1 = f3 | f4<<9 | dst<<18``;
  ``P[dst] = (((RCT[r]*O[f1])*O[f2])*O[f3])*O[f4]`` with O = V | FIX | literals | 1.0
* sum stream, 16-bit words in groups of GS: GS-1 terms ``idx<<3 | minus`` followed by an
  output word ``idx<<3 | store | diag<<1``; the running sum continues through groups
  whose output word has store = 0
* elimination ops, four 16-bit words ``d<<3|flags, c<<3, a<<3, b<<3``:
  ``G[d] = G[c] - G[a]*G[b]``; flags 1 = reciprocal pivot (``G[d] = -1/G[d]``),
  2 = division (strict build: ``G[d] = G[a]/G[b]``), 4 = block barrier after this op
* solve frames, eight 16-bit words ``row<<3|BEGIN|STORE<<1|SCALE<<2``,
  ``gp<<3|WSYNC|CSYNC<<1|PARTIAL<<2``, six entries ``col<<3|valid``
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import mech as mechmod

CHUNK = 16         # 16-bit stream words per thread and pipeline stage (32 bytes)
FUN_GS, JAC_GS = 8, 4
F_BEGIN, F_STORE, F_SCALE = 1, 2, 4
F_WSYNC, F_CSYNC, F_PARTIAL = 1, 2, 4
H_RECIP, H_DIV, H_SYNC = 1, 2, 4
NENT = 6           # entries per solve frame


def _lpt(weights, nbins, load=None):
    """Longest-processing-time assignment: returns bins[i] for every item."""
    order = sorted(range(len(weights)), key=lambda i: -weights[i])
    load = [0] * nbins if load is None else load
    out = [0] * len(weights)
    for i in order:
        b = min(range(nbins), key=lambda k: (load[k], k))
        out[i] = b
        load[b] += weights[i]
    return out, load


class Plan:
    def __init__(self, m, T, strict=False, split=24):
        self.m = m
        self.T = T
        self.strict = strict
        self.split = None if strict else split     # fast build: long tail rows of the forward solve in parts
        n = self.n = m.nvar
        h = self.h = n - T
        assert T % 32 == 0 and 0 < T <= n
        self.R = T // 32            # tail rows per lane
        self.W = T // 32            # warps per cell (each owns 32 tail columns)
        self.NT = 32 * self.W
        pos = self.pos = m.pos
        self.low = [[] for _ in range(n)]
        self.up = [[] for _ in range(n)]
        for (i, c) in pos:
            if c < i:
                self.low[i].append(c)
            elif c > i:
                self.up[i].append(c)
        for l in self.low:
            l.sort()
        for u in self.up:
            u.sort()
        # ---- G: head diagonals D[0..h), then every off-diagonal entry outside the tail block
        self.gidx = np.full(m.lu_nonzero, -1, dtype=np.int64)
        for k in range(h):
            self.gidx[int(m.diag[k])] = k
        g = h
        for nz in range(m.lu_nonzero):
            r, c = int(m.row_of[nz]), int(m.icol[nz])
            if (r >= h and c >= h) or (r == c):
                continue
            self.gidx[nz] = g
            g += 1
        self.NG = g
        self.ZERO = g               # slot that always holds 0.0
        self.NGP = (g + 2) & ~1
        self.fill = [nz for nz in range(m.lu_nonzero) if not m.jvs[nz] and self.gidx[nz] >= h]
        # ---- operand space of the rate products: V | FIX | literals | 1.0 -------------
        self.lits = list(m.coef_literals)
        self.nlit = len(self.lits)
        self.nc = m.nfix + self.nlit + 1
        self.ONE = n + m.nfix + self.nlit
        assert n + self.nc <= 512 and m.nreact < 2048
        self._layout()
        self._build_fun()
        self._build_jac()
        self._build_hops()
        self._build_ht()
        self._build_solve()

    # ---- shared-memory map (doubles) -----------------------------------------------------------
    def _scaled_pairs(self, tab):
        pairs = []
        for t in tab:
            for s, c, i in t:
                if c is not None and (int(i), c) not in pairs:
                    pairs.append((int(i), c))
        return pairs

    def _layout(self):
        m, n = self.m, self.n
        self.fun_scaled = self._scaled_pairs(m.vdot)
        self.jac_scaled = self._scaled_pairs(m.jvs)
        o = 0
        self.O_G = o; o += self.NGP
        self.O_Y = o; o += n
        self.O_CY = o; o += self.nc
        o += o & 1
        self.O_T1 = o; o += n
        self.O_CT = o; o += self.nc
        # every K vector is followed by a slot that holds 0.0 while the vector is being solved for: padded
        # entries of the substitution frames point there (x = 0), so the frames run without branches
        self.O_K1 = o; o += n + 1
        o += o & 1
        self.O_K2 = o; o += n + 1      # 16-byte aligned: RCONST is staged here by a bulk copy
        self.O_K3 = o; o += n + 1
        self.O_EX = o
        # Fun: [K2..EX) holds RCT/A (in place, shifted by <= 1), the scaled copies, a zero and a dump slot
        self.fun_nscr = m.nreact + 1 + len(self.fun_scaled) + 2
        need_fun = self.fun_nscr - (self.O_EX - self.O_K2)
        # Jac: B in [T1..EX) (the scaled copies go to fill-in slots of G, RCT is staged in G)
        late = [int(self.gidx[nz]) for nz in self.fill if self.gidx[nz] >= m.nreact + 2]
        self.jac_in_fill = len(late) >= len(self.jac_scaled)       # else: dedicated scratch after B
        need_jac = m.bdim + 1 + (0 if self.jac_in_fill else len(self.jac_scaled)) - (self.O_EX - self.O_T1)
        self.lbuf = 2 * self.T if self.W > 1 else 0
        self.ubuf = 2 * 32 * self.W
        need_lu = self.lbuf + self.ubuf - (self.O_EX - self.O_K2)
        need_xp = self.T
        self.nex = max(0, need_jac, need_fun, need_lu, need_xp)
        o += self.nex
        o += o & 1
        self.O_MISC = o; o += 12 + self.W                           # integrator state, reduction scratch, RCONST mbarrier
        o += o & 1
        self.smem_doubles = o                                       # one cell slot; the block adds the stream pipeline
        assert o < 8192
        assert m.nreact + 2 <= self.NG, "RCONST staging for the Jacobian does not fit in G"

    # ------------------------------------------------------------------------------
    def opref(self, f):
        kind, v = f
        if kind == "V":
            return int(v)
        if kind == "F":
            return self.n + int(v)
        if kind == "N":
            return self.n + self.m.nfix + self.lits.index(v)
        raise ValueError(kind)

    def litref(self, c):
        return self.n + self.m.nfix + self.lits.index(c)

    def lit_values(self, f32):
        return np.asarray([self.m.literal_value(c, f32) for c in self.lits])

    def const_block(self, fix, f32):
        """Values of the constant operands that follow V in shared memory."""
        return np.concatenate([np.asarray(fix, dtype=np.float64), self.lit_values(f32), [1.0]])

    def _pack16(self, per_thread):
        """per_thread[t] = list of 16-bit words.  Returns (flat uint16 array, nchunk): every thread's stream is
        padded to the same multiple of CHUNK words and stored as [chunk][w][half][lane][8]: a chunk (W KiB) is
        what one stage of the block's stream pipeline holds; warp w reads its 1 KiB piece, a thread's two
        16-byte items sit at [half][lane] (conflict-free 128-bit shared loads)."""
        L = max(len(x) for x in per_thread)
        L = max(CHUNK, (L + CHUNK - 1) // CHUNK * CHUNK)
        a = np.zeros((self.NT, L), dtype=np.uint16)
        for t in range(self.NT):
            ws = per_thread[t]
            a[t, :len(ws)] = ws
        a = a.reshape(self.W, 32, L // CHUNK, 2, 8).transpose(2, 0, 3, 1, 4)
        return np.ascontiguousarray(a).reshape(-1), L // CHUNK

    def unpack16(self, flat, nchunk):
        a = flat.reshape(nchunk, self.W, 2, 32, 8).transpose(1, 3, 0, 2, 4).reshape(self.NT, -1)
        return [a[t] for t in range(self.NT)]

    # ---- rate products -----------------------------------------------------------------
    def _prod_words(self, r, refs, dst):
        refs = list(refs) + [self.ONE] * (4 - len(refs))
        assert len(refs) == 4 and dst < 8192
        w0 = r | (refs[0] << 11) | (refs[1] << 20)
        w1 = refs[2] | (refs[3] << 9) | (dst << 18)
        return [w0 & 0xFFFF, w0 >> 16, w1 & 0xFFFF, w1 >> 16]

    def _prod_stream(self, ops, dump):
        """ops = [(r, refs, dst)] -> per-thread round robin, padded with ops that write the dump slot."""
        per = [[] for _ in range(self.NT)]
        for i, (r, refs, dst) in enumerate(ops):
            per[i % self.NT] += self._prod_words(r, refs, dst)
        L = max(len(x) for x in per)
        L = (L + CHUNK - 1) // CHUNK * CHUNK
        for t in range(self.NT):
            while len(per[t]) < L:
                per[t] += self._prod_words(0, [], dump)
        return self._pack16(per)

    def _sum_stream(self, items, GS):
        """items = [(terms as (idx, minus), out idx, diag flag)] -> grouped per-thread streams."""
        groups = []
        for terms, oidx, dflag in items:
            g = []
            k = 0
            ng = max(1, (len(terms) + GS - 2) // (GS - 1))
            for gi in range(ng):
                part = terms[gi * (GS - 1):(gi + 1) * (GS - 1)]
                ws = [(int(i) << 3) | (1 if neg else 0) for i, neg in part]
                ws += [(self.szero << 3)] * (GS - 1 - len(part))
                last = gi == ng - 1
                ws.append(((int(oidx) << 3) | 1 | (2 if dflag else 0)) if last else 0)
                g += ws
            groups.append(g)
        who, _ = _lpt([len(g) for g in groups], self.NT)
        per = [[] for _ in range(self.NT)]
        for k in sorted(range(len(groups)), key=lambda k: -len(groups[k])):
            per[who[k]] += groups[k]
        # pad with empty groups
        L = max(len(x) for x in per)
        L = (L + CHUNK - 1) // CHUNK * CHUNK
        for t in range(self.NT):
            while len(per[t]) < L:
                per[t] += [(self.szero << 3)] * (GS - 1) + [0]
        return self._pack16(per)

    # ---- Fun: Vdot(i) = sum of +-c*A(r), reference term order -----------------------------
    def _build_fun(self):
        """Indices relative to the staging base (RCT row, then scaled copies, zero, dump)."""
        m = self.m
        nr = m.nreact
        sc = {p: nr + k for k, p in enumerate(self.fun_scaled)}
        self.fun_zero = nr + len(sc)
        self.fun_dump = self.fun_zero + 1
        # pass 1: the scaled copies c*A(r) (they read RCT[r] before pass 2 overwrites it with A(r))
        ops = []
        for (r, c), dst in sc.items():
            facs = m.reactions[r]
            ops.append((r, [self.opref(f) for f in facs[1:]] + [self.ONE] * (4 - len(facs)) + [self.litref(c)], dst))
            assert len(ops[-1][1]) == 4
        self.funs_stream, self.funs_nchunk = self._prod_stream(ops, self.fun_dump)
        ops = []
        for r, facs in enumerate(m.reactions):
            assert facs[0][0] == "R" and int(facs[0][1]) == r and len(facs) <= 4
            ops.append((r, [self.opref(f) for f in facs[1:]], r))
        self.funp_stream, self.funp_nchunk = self._prod_stream(ops, self.fun_dump)
        self.szero = self.fun_zero
        items = []
        for i in range(self.n):
            terms = [(int(a) if c is None else sc[(int(a), c)], s < 0) for s, c, a in m.vdot[i]]
            items.append((terms, i, False))
        self.fun_stream, self.fun_nchunk = self._sum_stream(items, FUN_GS)

    # ---- Jac_SP fused with Ghimj = -Jac0 ; diag += 1/(H*gamma) ----------------------------
    def _build_jac(self):
        """Absolute shared-memory indices.  B(m) at O_T1 + m; scaled copies c*B(m) in fill-in slots of G
        beyond the RCT staging area; tail entries are staged in G[0..) and picked up by their owners."""
        m, h = self.m, self.h
        nr = m.nreact
        if self.jac_in_fill:
            late = [int(self.gidx[nz]) for nz in self.fill if self.gidx[nz] >= nr + 2]
            sc = {p: late[k] for k, p in enumerate(self.jac_scaled)}
            self.jac_late = sorted(sc.values())
        else:
            sc = {p: self.O_T1 + m.bdim + 1 + k for k, p in enumerate(self.jac_scaled)}
            self.jac_late = []
        self.jac_dump = self.O_T1 + m.bdim
        Bd = {int(b): facs for b, facs in m.B}
        ops = []
        for b in range(m.bdim):
            if b in Bd:
                facs = Bd[b]
                ops.append((int(facs[0][1]), [self.opref(f) for f in facs[1:]], self.O_T1 + b))
        for (b, c), dst in sc.items():
            facs = Bd[b]
            ops.append((int(facs[0][1]), [self.opref(f) for f in facs[1:]] + [self.ONE] * (4 - len(facs)) + [self.litref(c)], dst))
        self.jacp_stream, self.jacp_nchunk = self._prod_stream(ops, self.jac_dump)
        self.szero = self.ZERO

        def terms_of(nz):
            return [(self.O_T1 + int(b) if c is None else sc[(int(b), c)], s < 0) for s, c, b in m.jvs[nz]]
        # tail entries that Jac_SP assigns: staging list ordered by register (q, s), then owner thread
        self.pick_mask = np.zeros((self.W, self.R, 32), dtype=np.uint32)
        self.pick_base = np.zeros((self.W, self.R, 32), dtype=np.uint16)
        items = []
        for q in range(self.R):
            for s in range(32):
                for w in range(self.W):
                    self.pick_base[w, q, s] = len(items)
                    for p in range(32):
                        e = (h + p + 32 * q, h + 32 * w + s)
                        if e in self.pos and m.jvs[self.pos[e]]:
                            self.pick_mask[w, q, s] |= np.uint32(1 << p)
                            items.append((terms_of(self.pos[e]), len(items), False))
        self.ntstage = len(items)
        assert self.ntstage <= self.NG
        self.jt_stream, self.jt_nchunk = self._sum_stream(items, JAC_GS)
        # head entries: every slot of G except the late ones (those hold scaled products until the end)
        lateset = set(self.jac_late)
        items = []
        fills = []
        for nz in range(m.lu_nonzero):
            gi = int(self.gidx[nz])
            if gi >= 0 and gi not in lateset:
                if m.jvs[nz] or gi < h:
                    items.append((terms_of(nz), gi, gi < h))
                else:
                    fills.append(gi)            # fill-in: Jac_SP leaves 0, Ghimj = -0.0
        self.jh_stream, self.jh_nchunk = self._sum_stream(items, JAC_GS)
        self.jfill_stream, self.jfill_nchunk = self._fill_stream(fills)
        self.jlate_stream, self.jlate_nchunk = self._fill_stream(self.jac_late)

    def _fill_stream(self, idx):
        per = [[] for _ in range(self.NT)]
        for i, g in enumerate(idx):
            per[i % self.NT].append(int(g) << 3)
        L = max(CHUNK, (max(len(x) for x in per) + CHUNK - 1) // CHUNK * CHUNK)
        for t in range(self.NT):
            per[t] += [self.jac_dump << 3] * (L - len(per[t]))
        return self._pack16(per)

    # ---- elimination of everything that lives in G (KppDecomp order, wave-scheduled) -------
    def head_ops(self):
        """Yield (flags, d, c, a, b) in KppDecomp's program order, restricted to destinations in G:
        G[d] = G[c] - G[a]*G[b].  Fast build: head diagonals end up as -1/pivot, a multiplier is
        0 - L*(-1/pivot); strict build: true division."""
        m, h, g, pos, Z = self.m, self.h, self.gidx, self.pos, self.ZERO
        for k in range(self.n):
            for kk in range(int(m.crow[k]), int(m.diag[k])):
                j = int(m.icol[kk])
                if j >= h:
                    break
                if self.strict:
                    yield (H_DIV, int(g[kk]), Z, int(g[kk]), j)
                else:
                    yield (0, int(g[kk]), Z, int(g[kk]), j)
                for jj in range(int(m.diag[j]) + 1, int(m.crow[j + 1])):
                    c = int(m.icol[jj])
                    if k >= h and c >= h:
                        continue
                    d = int(g[pos[(k, c)]])
                    yield (0, d, d, int(g[kk]), int(g[jj]))
            if k < h and not self.strict:
                yield (H_RECIP, k, k, k, k)

    def _build_hops(self):
        depth = {}
        waves = {}
        for op in self.head_ops():
            fl, d, c, a, b = op
            if fl == H_RECIP:
                dd = depth.get(d, 0) + 1
            else:
                dd = max(depth.get(c, 0), depth.get(a, 0), depth.get(b, 0)) + 1
            depth[d] = dd
            waves.setdefault(dd, []).append(op)
        per = [[] for _ in range(self.NT)]
        self.hop_waves = []
        Z = self.ZERO
        for dd in sorted(waves):
            ops = sorted(waves[dd], key=lambda o: -(o[0] & H_RECIP))   # reciprocals share the first slots
            self.hop_waves.append(ops)
            nslot = (len(ops) + self.NT - 1) // self.NT
            for t in range(self.NT):
                mine = ops[t::self.NT]
                mine = mine + [(0, Z, Z, Z, Z)] * (nslot - len(mine))
                for i, (fl, d, c, a, b) in enumerate(mine):
                    if i == nslot - 1:
                        fl |= H_SYNC
                    per[t] += [(d << 3) | fl, c << 3, a << 3, b << 3]
        L = max(len(x) for x in per)
        L = (L + CHUNK - 1) // CHUNK * CHUNK
        for t in range(self.NT):
            while len(per[t]) < L:
                per[t] += [Z << 3, Z << 3, Z << 3, Z << 3]
        self.hop_stream, self.hop_nchunk = self._pack16(per)

    # ---- head pivots acting on the register tail --------------------------------------------
    def _build_ht(self):
        """For every head pivot j that has entries in tail rows: lane masks per row slot q (which
        of the rows h+p+32q hold L(row, j): their multipliers are consumed in CSR order through a
        running pointer per row) and the tail columns of U(j, :)."""
        m, h, n = self.m, self.h, self.n
        self.ht = []            # (j, masks[q], [(c - h, gidx of U(j,c))])
        for j in range(h):
            masks = [0] * self.R
            for i in range(h, n):
                if (i, j) in self.pos:
                    masks[(i - h) // 32] |= 1 << ((i - h) % 32)
            cols = [(c - h, int(self.gidx[self.pos[(j, c)]])) for c in self.up[j] if c >= h]
            if any(masks):
                self.ht.append((j, masks, cols))
        # first L entry of every tail row
        self.rowbase = np.asarray([int(self.gidx[int(m.crow[h + r])]) if self.low[h + r] and self.low[h + r][0] < h
                                   else self.ZERO for r in range(self.T)], dtype=np.uint16)

    # ---- triangular solves: the part outside the tail block ---------------------------------
    def _frames(self, row, gp, cols, begin, store, scale, partial=False):
        """Frames of one piece: acc (= X[row] | 0) -= sum G[gp+i]*X[cols[i]]; optionally stored."""
        out = []
        nf = max(1, (len(cols) + NENT - 1) // NENT)
        for f in range(nf):
            part = cols[f * NENT:(f + 1) * NENT]
            h0 = row << 3
            if f == 0 and begin:
                h0 |= F_BEGIN
            if f == nf - 1:
                if store:
                    h0 |= F_STORE
                if scale:
                    h0 |= F_SCALE
            h1 = ((gp + f * NENT) << 3) | (F_PARTIAL if partial else 0)
            out.append([h0, h1] + [(c << 3) | 1 for c in part] + [self.n << 3] * (NENT - len(part)))
        return out

    def _schedule(self, levels, warp0_only):
        """levels = list of lists of pieces (piece = list of frames).  Returns per-thread word lists.
        Levels listed in warp0_only are run by warp 0 alone (the other warps idle through empty frames) and
        are closed by a warp barrier when the next level is warp 0's too; all other levels are spread over
        all threads and closed by a barrier of the cell's threads.  Every thread's stream has the same
        length at every barrier, so the warps of a cell stay within one chunk of each other."""
        per = [[] for _ in range(self.NT)]
        nul = [0, self.ZERO << 3] + [self.n << 3] * NENT
        allt = list(range(self.NT))
        for li, pieces in enumerate(levels):
            w0 = li in warp0_only
            ts = list(range(32)) if w0 else allt
            who, _ = _lpt([len(p) for p in pieces], len(ts))
            for k in sorted(range(len(pieces)), key=lambda k: -len(pieces[k])):
                for fr in pieces[k]:
                    per[ts[who[k]]] += fr
            L = max(len(per[t]) for t in allt)
            for t in allt:
                while len(per[t]) < L:
                    per[t] += nul
            flag = F_WSYNC if (w0 and (li + 1) in warp0_only) else F_CSYNC
            for t in allt:
                per[t][len(per[t]) - 8 + 1] |= flag
        return per

    def _build_solve(self):
        m, h, n, g, pos = self.m, self.h, self.n, self.gidx, self.pos
        # ---- forward: x_k -= sum_{j<k, j<h} L(k,j) x_j in ascending j.  Head rows by dependency level;
        # then the tail rows (fast build: rows longer than `split` in parts whose partial sums are added
        # when the tail sweep loads them).
        lvl = [0] * n
        levels = {}
        for k in range(h):
            cols = self.low[k]
            if not cols:
                continue
            lvl[k] = 1 + max(lvl[j] for j in cols)
            levels.setdefault(lvl[k], []).append(
                self._frames(k, int(g[pos[(k, cols[0])]]), cols, True, True, False))
        head_levels = [levels[l] for l in sorted(levels)]
        tail = []
        self.fwd_partial = np.zeros(self.T, dtype=np.int32)       # number of extra partial sums per tail row
        for k in range(h, n):
            cols = [j for j in self.low[k] if j < h]
            if not cols:
                continue
            gp = int(g[pos[(k, cols[0])]])
            if self.split and len(cols) > self.split:
                half = (len(cols) + 1) // 2
                tail.append(self._frames(k, gp, cols[:half], True, True, False))
                tail.append(self._frames(k - h, gp + half, cols[half:], True, True, False, partial=True))
                self.fwd_partial[k - h] = 1
            else:
                tail.append(self._frames(k, gp, cols, True, True, False))
        lv = head_levels + ([tail] if tail else [])
        small = set(i for i, p in enumerate(head_levels) if len(p) <= 40) if self.W > 1 else set()
        per = self._schedule(lv, small)
        self.fwd_levels = lv
        self.fwd_stream, self.fwd_nchunk = self._pack16(per)
        # ---- backward: head rows (the tail block is solved first)
        #   strict: reference order (ascending columns, then the division)
        #   fast:   tail columns first (known once the tail is solved), then head columns by level
        blvl = [0] * n
        levels = {}
        for k in range(h - 1, -1, -1):
            hc = [c for c in self.up[k] if c < h]
            tc = [c for c in self.up[k] if c >= h]
            first = int(g[pos[(k, self.up[k][0])]]) if self.up[k] else self.ZERO
            if self.strict:
                blvl[k] = 1 + max([blvl[c] for c in hc], default=0)
                levels.setdefault(blvl[k], []).append(self._frames(k, first, hc + tc, True, True, True))
            elif hc:
                blvl[k] = 1 + max(blvl[c] for c in hc)
                if tc:
                    levels.setdefault(0, []).append(
                        self._frames(k, int(g[pos[(k, tc[0])]]), tc, True, True, False))
                levels.setdefault(blvl[k], []).append(self._frames(k, first, hc, True, True, True))
            else:
                blvl[k] = 0
                levels.setdefault(0, []).append(self._frames(k, first, tc, True, True, True))
        lv = [levels[l] for l in sorted(levels)]
        small = set(i for i, p in enumerate(lv) if i > 0 and len(p) <= 40) if self.W > 1 else set()
        per = self._schedule(lv, small)
        self.bwd_levels = lv
        self.bwd_stream, self.bwd_nchunk = self._pack16(per)

    # =============================================================================================
    # CPU emulation (numpy, IEEE double, no fused multiply-add): same order as the kernel
    # =============================================================================================
    def new_smem(self):
        S = np.zeros(self.smem_doubles)          # the kernel zero-fills every slot before its first cell
        return S

    def emu_set_consts(self, S, fix, f32):
        cb = self.const_block(fix, f32)
        S[self.O_CY:self.O_CY + self.nc] = cb
        S[self.O_CT:self.O_CT + self.nc] = cb

    def emu_products(self, flat, nchunk, S, rbase, obase, dbase):
        P = self.unpack16(flat, nchunk)
        for t in range(self.NT):
            ws = P[t]
            for i in range(0, len(ws), 4):
                w0 = int(ws[i]) | (int(ws[i + 1]) << 16)
                w1 = int(ws[i + 2]) | (int(ws[i + 3]) << 16)
                r, f1, f2 = w0 & 0x7FF, (w0 >> 11) & 0x1FF, (w0 >> 20) & 0x1FF
                f3, f4, dst = w1 & 0x1FF, (w1 >> 9) & 0x1FF, w1 >> 18
                S[dbase + dst] = (((S[rbase + r] * S[obase + f1]) * S[obase + f2]) * S[obase + f3]) * S[obase + f4]

    def emu_sums(self, flat, nchunk, S, GS, sbase, obase, neg=False, ghinv=0.0):
        P = self.unpack16(flat, nchunk)
        writes = []
        sing = False
        for t in range(self.NT):
            acc = 0.0
            ws = P[t]
            for i in range(0, len(ws), GS):
                for k in range(GS - 1):
                    w = int(ws[i + k])
                    v = S[sbase + (w >> 3)]
                    acc = acc - v if (w & 1) else acc + v
                w = int(ws[i + GS - 1])
                if w & 1:
                    v = (-0.0 - acc) if neg else acc
                    if w & 2:
                        v = ghinv - acc
                        sing |= (v == 0.0)
                    writes.append((obase + (w >> 3), v))
                    acc = 0.0
        for i, v in writes:                 # outputs never overlap the sources of the same pass
            S[i] = v
        return sing

    def emu_fun(self, S, vbase, outbase, rct):
        """S[outbase + i] = Vdot(i) for V = S[vbase..]; scratch = [K2..EX)."""
        nr = self.m.nreact
        rb = self.O_K2 + 1                                    # any shift (0 or 1) the bulk copy chooses
        S[rb:rb + nr] = rct
        S[rb + self.fun_zero] = 0.0
        self.emu_products(self.funs_stream, self.funs_nchunk, S, rb, vbase, rb)
        self.emu_products(self.funp_stream, self.funp_nchunk, S, rb, vbase, rb)
        self.emu_sums(self.fun_stream, self.fun_nchunk, S, FUN_GS, rb, outbase)

    def emu_jacprep(self, S, rct, ghinv):
        """G and the tail registers a[t][q][s] from V = S[O_Y..]."""
        nr = self.m.nreact
        rb = 1
        S[rb:rb + nr] = rct                                   # RCT staged in G
        self.emu_products(self.jacp_stream, self.jacp_nchunk, S, rb, self.O_Y, 0)
        S[self.ZERO] = 0.0
        self.emu_sums(self.jt_stream, self.jt_nchunk, S, JAC_GS, 0, 0, neg=True)
        a = np.zeros((self.NT, self.R, 32))
        sing = False
        for t in range(self.NT):
            w, p = t >> 5, t & 31
            for q in range(self.R):
                for s in range(32):
                    mk = int(self.pick_mask[w, q, s])
                    v = -0.0
                    if (mk >> p) & 1:
                        v = S[int(self.pick_base[w, q, s]) + bin(mk & ((1 << p) - 1)).count("1")]
                    if q == w and s == p:
                        v = v + ghinv
                        sing |= (v == 0.0)
                    a[t, q, s] = v
        for flat, nch in ((self.jfill_stream, self.jfill_nchunk),):
            for ws in self.unpack16(flat, nch):
                for w in ws:
                    S[int(w) >> 3] = -0.0
        sing |= self.emu_sums(self.jh_stream, self.jh_nchunk, S, JAC_GS, 0, 0, neg=True, ghinv=ghinv)
        for ws in self.unpack16(self.jlate_stream, self.jlate_nchunk):
            for w in ws:
                S[int(w) >> 3] = -0.0
        S[self.T1dead():self.O_EX + self.nex] = np.nan        # B scratch is dead now (K1.. get rewritten)
        self.emu_set_consts_ct(S)
        return a, sing

    def T1dead(self):
        return self.O_T1

    def emu_set_consts_ct(self, S):
        S[self.O_CT:self.O_CT + self.nc] = S[self.O_CY:self.O_CY + self.nc]

    def emu_hops(self, S):
        P = self.unpack16(self.hop_stream, self.hop_nchunk)
        nslot = len(P[0]) // 4
        for i in range(nslot):
            old = S.copy()          # no hazards between the ops of one wave
            for t in range(self.NT):
                d, c, a, b = (int(P[t][4 * i + k]) for k in range(4))
                fl, d, c, a, b = d & 7, d >> 3, c >> 3, a >> 3, b >> 3
                assert (fl & H_SYNC) == (int(P[0][4 * i]) & H_SYNC)
                if fl & H_RECIP:
                    S[d] = -1.0 / old[d]
                elif fl & H_DIV:
                    S[d] = old[a] / old[b]
                else:
                    S[d] = old[c] - old[a] * old[b]

    def emu_ht(self, S, a):
        gp = np.zeros((32, self.R), dtype=np.int64)
        for p in range(32):
            for q in range(self.R):
                gp[p, q] = self.rowbase[p + 32 * q]
        for (j, masks, cols) in self.ht:
            l = np.zeros((32, self.R))
            for p in range(32):
                for q in range(self.R):
                    if (masks[q] >> p) & 1:
                        l[p, q] = S[gp[p, q]]
                        gp[p, q] += 1
            for t in range(self.NT):
                w, p = t >> 5, t & 31
                for q in range(self.R):
                    if masks[q] == 0:
                        continue
                    for (c, gu) in cols:
                        if c // 32 == w:
                            a[t, q, c - 32 * w] = a[t, q, c - 32 * w] - l[p, q] * S[gu]

    def emu_tail_lu(self, a):
        """Right-looking elimination of the register tail, pivot by pivot."""
        T = self.T

        def A(r, c):
            return (32 * (c // 32) + (r % 32), r // 32, c % 32)
        for k in range(T):
            d = a[A(k, k)]
            if not self.strict:
                rp = 1.0 / d
                a[A(k, k)] = rp
            l = np.zeros(T)
            for r in range(k + 1, T):
                l[r] = a[A(r, k)] / d if self.strict else a[A(r, k)] * rp
                a[A(r, k)] = l[r]
            for c in range(k + 1, T):
                u = a[A(k, c)]
                for r in range(k + 1, T):
                    a[A(r, c)] = a[A(r, c)] - l[r] * u

    def emu_decomp(self, S, a):
        self.emu_hops(S)
        self.emu_ht(S, a)
        self.emu_tail_lu(a)

    def emu_frames(self, flat, nchunk, S, xb, xpb):
        P = self.unpack16(flat, nchunk)
        pos_ = [0] * self.NT
        acc = [0.0] * self.NT
        gp = [0] * self.NT
        done = [False] * self.W

        def run_warp(w):
            """advance warp w up to and including its next barrier frame; returns the flag"""
            flag = None
            while flag is None:
                writes = []
                for p in range(32):
                    t = 32 * w + p
                    ws = P[t]
                    i = pos_[t]
                    h0, h1 = int(ws[i]), int(ws[i + 1])
                    row = h0 >> 3
                    partial = h1 & F_PARTIAL
                    if h0 & F_BEGIN:
                        acc[t] = 0.0 if partial else S[xb + row]
                        gp[t] = h1 >> 3
                    for k in range(NENT):
                        e = int(ws[i + 2 + k])
                        if e & 1:
                            acc[t] = acc[t] - S[gp[t] + k] * S[xb + (e >> 3)]
                        elif h0 & (F_BEGIN | F_STORE) or (h0 >> 3):
                            # the kernel has no branch here: a padded entry multiplies whatever follows the row in G
                            # by the zero slot behind the vector - that operand must be a finite number
                            assert (e >> 3) == self.n and S[xb + self.n] == 0.0
                            assert np.isfinite(S[gp[t] + k]), "padded entry reads a non-finite operand"
                            acc[t] = acc[t] - S[gp[t] + k] * S[xb + self.n]
                    gp[t] += NENT
                    if h0 & F_STORE:
                        v = acc[t]
                        if h0 & F_SCALE:
                            v = v / S[row] if self.strict else -(v * S[row])
                        writes.append(((xpb if partial else xb) + row, v))
                    pos_[t] = i + 8
                    f = h1 & 3
                    if p == 0:
                        f0 = f
                    assert f == f0
                for i, v in writes:
                    S[i] = v
                if f0:
                    flag = f0
                if pos_[32 * w] >= len(P[32 * w]) and flag is None:
                    flag = -1
            return flag
        # warps run independently between block barriers
        while True:
            flags = []
            for w in range(self.W):
                while True:
                    if pos_[32 * w] >= len(P[32 * w]):
                        flags.append(-1)
                        break
                    f = run_warp(w)
                    if f == F_WSYNC:
                        continue
                    flags.append(f)
                    break
            assert len(set(flags)) == 1, flags
            if flags[0] == -1:
                break

    def emu_solve(self, S, a, xb):
        T, h = self.T, self.h
        xpb = self.O_EX

        def A(r, c):
            return (32 * (c // 32) + (r % 32), r // 32, c % 32)
        S[xpb:xpb + T] = 0.0
        S[xb + self.n] = 0.0                     # the zero slot behind the vector (kernel: set with the right-hand side)
        self.emu_frames(self.fwd_stream, self.fwd_nchunk, S, xb, xpb)
        X = S[xb:xb + self.n]
        for r in range(T):
            if self.fwd_partial[r]:
                X[h + r] = X[h + r] + S[xpb + r]
        for c in range(T):                       # forward, column by column (ascending)
            for r in range(c + 1, T):
                X[h + r] = X[h + r] - a[A(r, c)] * X[h + c]
        if self.strict:
            for r in range(T - 1, -1, -1):       # reference order: ascending columns, then divide
                acc = X[h + r]
                for c in range(r + 1, T):
                    acc = acc - a[A(r, c)] * X[h + c]
                X[h + r] = acc / a[A(r, r)]
        else:
            for c in range(T - 1, -1, -1):       # column by column (descending), reciprocal pivots
                X[h + c] = X[h + c] * a[A(c, c)]
                for r in range(c):
                    X[h + r] = X[h + r] - a[A(r, c)] * X[h + c]
        self.emu_frames(self.bwd_stream, self.bwd_nchunk, S, xb, xpb)

    def emu_to_csr(self, S, a, undo_recip=False):
        """LU values in the reference's storage order."""
        m = self.m
        out = np.zeros(m.lu_nonzero)
        for nz in range(m.lu_nonzero):
            if self.gidx[nz] >= 0:
                out[nz] = S[self.gidx[nz]]
            else:
                r, c = int(m.row_of[nz]) - self.h, int(m.icol[nz]) - self.h
                out[nz] = a[32 * (c // 32) + (r % 32), r // 32, c % 32]
        return out


# =================================================================================================
# CUDA emission
# =================================================================================================
STREAMS = ("funs", "funp", "fun", "jacp", "jt", "jfill", "jh", "jlate", "hop", "fwd", "bwd")


def table_blob(p):
    """All streams of a plan in one uint16 array; returns (blob, desc) with desc[name] = (first chunk, number of
    chunks); a chunk = W KiB = CHUNK words for each of the cell's threads."""
    parts = []
    desc = {}
    o = 0
    for name in STREAMS:
        flat = getattr(p, name + "_stream")
        nch = getattr(p, name + "_nchunk")
        assert len(flat) == nch * CHUNK * p.NT
        desc[name] = (o, nch)
        parts.append(flat)
        o += nch
    return np.concatenate(parts), desc


def emit_tables_cpp(p, sym):
    blob, _ = table_blob(p)
    lines = ["// GENERATED by mistra_b200/mechgen/onchip.py - instruction streams of the on-chip Ros3 kernel (%s, %s)."
             % (p.m.name, "strict" if p.strict else "fast"),
             "#include <cstddef>",
             "extern \"C\" const size_t %s_count = %d;" % (sym, len(blob)),
             "extern \"C\" const unsigned short %s[%d] = {" % (sym, len(blob))]
    for i in range(0, len(blob), 32):
        lines.append(",".join(str(int(v)) for v in blob[i:i + 32]) + ",")
    lines.append("};")
    return "\n".join(lines) + "\n"


class Emitter:
    def __init__(self, pf, ps):
        self.pf, self.ps = pf, ps
        self.L = []

    def w(self, s=""):
        self.L.append(s)

    def emit(self):
        p = self.pf
        m = p.m
        w = self.w
        x = m.suffix
        w("// GENERATED by mistra_b200/mechgen/onchip.py from mistra_b200/mech/%s.json - do not edit." % m.name)
        w("// On-chip Ros3 kernel of mechanism '%s': tail %d x %d in registers (%d warps x %d rows per lane)," % (m.name, p.T, p.T, p.W, p.R))
        w("// head (%d entries) in shared memory." % p.NG)
        w("#pragma once")
        w("namespace oc_%s {" % x)
        w("constexpr int NVAR = %d, NFIX = %d, NREACT = %d, LU_NONZERO = %d, BDIM = %d;" % (m.nvar, m.nfix, m.nreact, m.lu_nonzero, m.bdim))
        w("constexpr int T = %d, R = %d, W = %d, NT = %d, HEAD = %d, NG = %d, ZERO = %d, NGP = %d;" % (p.T, p.R, p.W, p.NT, p.h, p.NG, p.ZERO, p.NGP))
        w("constexpr int NC = %d, NLIT = %d;" % (p.nc, p.nlit))
        for k in ("O_G", "O_Y", "O_CY", "O_T1", "O_CT", "O_K1", "O_K2", "O_K3", "O_EX", "O_MISC"):
            w("constexpr int %s = %d;" % (k, getattr(p, k)))
            assert getattr(self.ps, k) == getattr(p, k)
        assert self.ps.smem_doubles == p.smem_doubles
        w("constexpr int NEX = %d, SMEM_DOUBLES = %d;" % (p.nex, p.smem_doubles))
        w("constexpr int FUN_ZERO = %d, FUN_DUMP = %d, JAC_DUMP = %d, NTSTAGE = %d;" % (p.fun_zero, p.fun_dump, p.jac_dump, p.ntstage))
        w("constexpr int LBUF = %d, UBUF = %d;   // elimination staging, doubles from O_K2" % (p.O_K2, p.O_K2 + p.lbuf))
        w("static const char *const coef_literals[NLIT] = {%s};" % ", ".join('"%s"' % c for c in p.lits))
        w("enum { ST_FUNS, ST_FUNP, ST_FUN, ST_JACP, ST_JT, ST_JFILL, ST_JH, ST_JLATE, ST_HOP, ST_FWD, ST_BWD, ST_COUNT };")
        for tag, pl in (("#ifdef KPP_STRICT", self.ps), ("#else", self.pf)):
            w(tag)
            blob, desc = table_blob(pl)
            w("constexpr unsigned TABLE_COUNT = %d;" % len(blob))
            w("// first chunk / number of chunks of every stream (a chunk = W KiB)")
            w("__constant__ unsigned c_st_off[ST_COUNT] = {%s};" % ", ".join(str(desc[n][0]) for n in STREAMS))
            w("__constant__ unsigned c_st_nchunk[ST_COUNT] = {%s};" % ", ".join(str(desc[n][1]) for n in STREAMS))
            for q in range(pl.R):
                w("#define FWD_PARTIAL_%d 0x%xu" % (q, sum((1 << pp) for pp in range(32) if pl.fwd_partial[pp + 32 * q])))
        w("#endif")
        w("__constant__ unsigned c_pick_mask[W][R][32] = {%s};" % ", ".join(
            "{" + ", ".join("{" + ",".join("0x%xu" % int(v) for v in p.pick_mask[ww, q]) + "}" for q in range(p.R)) + "}" for ww in range(p.W)))
        w("__constant__ unsigned short c_pick_base[W][R][32] = {%s};" % ", ".join(
            "{" + ", ".join("{" + ",".join(str(int(v)) for v in p.pick_base[ww, q]) + "}" for q in range(p.R)) + "}" for ww in range(p.W)))
        w("__device__ const unsigned short d_rowbase[T] = {%s};" % ",".join(str(int(v)) for v in p.rowbase))
        w("// keeps the compiler from hoisting a whole pivot row into registers (the tail already fills them)")
        w("#define OC_SCHED_FENCE() __syncwarp()")
        w("// barrier of one cell slot of the block (named barrier 1 + slot, the NT threads that own the cell)")
        w("__device__ __forceinline__ void oc_slot_sync(unsigned id) { asm volatile(\"bar.sync %0, %1;\" ::\"r\"(id), \"n\"(NT) : \"memory\"); }")
        w("#define OC_SLOT_SYNC() oc_slot_sync(slot_bar)")
        w("")
        self.emit_ht()
        self.emit_tail_lu()
        self.emit_sweeps()
        w("}  // namespace oc_%s" % x)
        return "\n".join(self.L) + "\n"

    # ---- head pivots -> register tail -----------------------------------------------------------
    def emit_ht(self):
        p, w = self.pf, self.w
        w("// Head pivots acting on the register tail: a[q][s] -= L(h+p+32q, j) * U(j, h+32w+s) for j ascending")
        w("// (the part of KppDecomp's row loop whose pivot lies in the head and whose target lies in the tail).")
        for ww in range(p.W):
            w("__device__ __forceinline__ void ht_update_w%d(double (&a)[R][32], const double *__restrict__ S, const unsigned p)" % ww)
            w("{")
            w("  const char *Sb = reinterpret_cast<const char *>(S);")
            for q in range(p.R):
                w("  unsigned g%d = (unsigned)d_rowbase[p + %d] * 8u;" % (q, 32 * q))
            for (j, masks, cols) in p.ht:
                mine = [(c - 32 * ww, gu) for (c, gu) in cols if c // 32 == ww]
                w("  {  // pivot %d" % j)
                for q in range(p.R):
                    if masks[q]:
                        if mine:
                            w("    double l%d = 0.0;" % q)
                            w("    if ((0x%xu >> p) & 1u) { l%d = *reinterpret_cast<const double *>(Sb + g%d); g%d += 8u; }" % (masks[q], q, q, q))
                        else:
                            w("    if ((0x%xu >> p) & 1u) g%d += 8u;" % (masks[q], q))
                for (s, gu) in mine:
                    w("    { const double u = S[%d];" % gu)
                    for q in range(p.R):
                        if masks[q]:
                            w("      a[%d][%d] -= l%d * u;" % (q, s, q))
                    w("    }")
                w("    OC_SCHED_FENCE();")
                w("  }")
            w("}")
            w("")

    # ---- elimination of the register tail ---------------------------------------------------------
    def emit_tail_lu(self):
        p, w = self.pf, self.w
        T, R, W = p.T, p.R, p.W
        w("// Right-looking elimination of the T x T register tail.  Pivot k = 32*kb + ks: warp kb owns column k")
        w("// (forms the multipliers, passes them to the warps on its right through LBUF), every warp w >= kb")
        w("// broadcasts its 32-column piece of pivot row k through UBUF and updates its rows below k.")
        w("#ifdef KPP_STRICT")
        w("#define TL_PIV(d) const double pv = (d)")
        w("#define TL_MUL(x) ((x) / pv)")
        w("#define TL_SETDIAG(reg) ")
        w("#else")
        w("#define TL_PIV(d) const double pv = 1.0 / (d)")
        w("#define TL_MUL(x) ((x) * pv)")
        w("#define TL_SETDIAG(reg) reg = pv")
        w("#endif")
        w("// (no __restrict__ on pointers into shared memory here: other threads write what this thread reads after a barrier)")
        w("__device__ __forceinline__ void tail_lu(double (&a)[R][32], double *S, const unsigned w, const unsigned p, const unsigned slot_bar)")
        w("{")
        w("  double *lbuf = S + LBUF;")
        w("  double *ubuf = S + UBUF + w * 64;")
        w("  double l[R];")
        for kb in range(W):
            last = kb == W - 1
            w("  // ---- pivot block %d" % kb)
            w("  if (w == %d) {" % kb)
            for ks in range(32):
                k = 32 * kb + ks
                par = k & 1
                w("    {  // pivot %d" % k)
                w("      TL_PIV(__shfl_sync(0xffffffffu, a[%d][%d], %d));" % (kb, ks, ks))
                w("      l[%d] = (p > %du) ? TL_MUL(a[%d][%d]) : 0.0;" % (kb, ks, kb, ks))
                w("      if (p > %du) a[%d][%d] = l[%d];" % (ks, kb, ks, kb))
                w("      if (p == %du) { TL_SETDIAG(a[%d][%d]); }" % (ks, kb, ks))
                for q in range(kb + 1, R):
                    w("      l[%d] = TL_MUL(a[%d][%d]); a[%d][%d] = l[%d];" % (q, q, ks, q, ks, q))
                if not last:
                    for q in range(kb, R):
                        w("      lbuf[%d + p] = l[%d];" % (par * T + 32 * q, q))
                    w("      OC_SLOT_SYNC();")
                if ks < 31:
                    s0 = ks + 1
                    w("      if (p == %du) {" % ks)
                    s = s0
                    if s & 1:
                        w("        ubuf[%d] = a[%d][%d];" % (par * 32 + s, kb, s))
                        s += 1
                    while s < 32:
                        w("        *reinterpret_cast<double2 *>(ubuf + %d) = make_double2(a[%d][%d], a[%d][%d]);" % (par * 32 + s, kb, s, kb, s + 1))
                        s += 2
                    w("      }")
                    w("      __syncwarp();")
                    s = s0
                    if s & 1:
                        w("      { const double u = ubuf[%d];" % (par * 32 + s))
                        for q in range(kb, R):
                            w("        a[%d][%d] -= l[%d] * u;" % (q, s, q))
                        w("      }")
                        s += 1
                    while s < 32:
                        w("      { const double2 u = *reinterpret_cast<const double2 *>(ubuf + %d);" % (par * 32 + s))
                        for q in range(kb, R):
                            w("        a[%d][%d] -= l[%d] * u.x; a[%d][%d] -= l[%d] * u.y;" % (q, s, q, q, s + 1, q))
                        w("      }")
                        if (s // 2) % 2 == 1:
                            w("      OC_SCHED_FENCE();")
                        s += 2
                w("    }")
            w("  }")
            if not last:
                w("  else if (w > %d) {" % kb)
                w("    for (unsigned ks = 0; ks < 32u; ++ks) {")
                w("      const unsigned par = ks & 1u;")
                w("      OC_SLOT_SYNC();")
                for q in range(kb, R):
                    w("      l[%d] = lbuf[par * %d + %d + p];" % (q, T, 32 * q))
                w("      if (p == ks) {")
                for s in range(0, 32, 2):
                    w("        *reinterpret_cast<double2 *>(ubuf + par * 32 + %d) = make_double2(a[%d][%d], a[%d][%d]);" % (s, kb, s, kb, s + 1))
                w("      }")
                w("      __syncwarp();")
                for s in range(0, 32, 2):
                    w("      { const double2 u = *reinterpret_cast<const double2 *>(ubuf + par * 32 + %d);" % s)
                    for q in range(kb, R):
                        w("        a[%d][%d] -= l[%d] * u.x; a[%d][%d] -= l[%d] * u.y;" % (q, s, q, q, s + 1, q))
                    w("      }")
                    if (s // 2) % 2 == 1:
                        w("      OC_SCHED_FENCE();")
                w("    }")
                w("  }")
                if kb > 0:
                    w("  else {")
                    w("    for (unsigned ks = 0; ks < 32u; ++ks) OC_SLOT_SYNC();")
                    w("  }")
        w("}")
        w("")

    # ---- triangular sweeps over the register tail ----------------------------------------------------
    def emit_sweeps(self):
        p, w = self.pf, self.w
        T, R, W, h = p.T, p.R, p.W, p.h
        w("// Forward substitution with the unit-lower tail block (ascending columns = KppSolve's order).")
        w("__device__ __forceinline__ void tail_forward(const double (&a)[R][32], double *X, const double *XP,")
        w("                                             const unsigned w, const unsigned p, const unsigned slot_bar)")
        w("{")
        w("  double x[R];")
        for wb in range(W):
            w("  if (w == %d) {" % wb)
            for q in range(wb, R):
                w("    x[%d] = X[%d + p];" % (q, h + 32 * q))
                if wb == 0:
                    w("    if ((FWD_PARTIAL_%d >> p) & 1u) x[%d] += XP[%d + p];" % (q, q, 32 * q))
            for s in range(32):
                w("    { const double xs = __shfl_sync(0xffffffffu, x[%d], %d);" % (wb, s))
                if s < 31:
                    w("      if (p > %du) x[%d] -= a[%d][%d] * xs;" % (s, wb, wb, s))
                for q in range(wb + 1, R):
                    w("      x[%d] -= a[%d][%d] * xs;" % (q, q, s))
                w("    }")
            for q in range(wb, R):
                w("    X[%d + p] = x[%d];" % (h + 32 * q, q))
            w("  }")
            w("  OC_SLOT_SYNC();")
        w("}")
        w("")
        w("#ifndef KPP_STRICT")
        w("// Backward substitution with the upper tail block, column by column from the right; the diagonal")
        w("// registers hold reciprocal pivots.")
        w("__device__ __forceinline__ void tail_backward(const double (&a)[R][32], double *X, const unsigned w, const unsigned p, const unsigned slot_bar)")
        w("{")
        w("  double x[R];")
        for wb in range(W - 1, -1, -1):
            w("  if (w == %d) {" % wb)
            for q in range(0, wb + 1):
                w("    x[%d] = X[%d + p];" % (q, h + 32 * q))
            for s in range(31, -1, -1):
                w("    { if (p == %du) x[%d] *= a[%d][%d];" % (s, wb, wb, s))
                w("      const double xs = __shfl_sync(0xffffffffu, x[%d], %d);" % (wb, s))
                if s > 0:
                    w("      if (p < %du) x[%d] -= a[%d][%d] * xs;" % (s, wb, wb, s))
                for q in range(0, wb):
                    w("      x[%d] -= a[%d][%d] * xs;" % (q, q, s))
                w("    }")
            for q in range(0, wb + 1):
                w("    X[%d + p] = x[%d];" % (h + 32 * q, q))
            w("  }")
            w("  OC_SLOT_SYNC();")
        w("}")
        w("#endif")
        w("")


def main(argv):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outdir = os.path.join(root, "mistra_b200", "csrc", "_gen")
    os.makedirs(outdir, exist_ok=True)
    for name in argv[1:] or ["gas", "aer"]:
        m = mechmod.load(name)
        T = TAIL[name]
        pf, ps = Plan(m, T, strict=False), Plan(m, T, strict=True)
        files = {
            "onchip_%s.cuh" % m.suffix: Emitter(pf, ps).emit(),
            "onchip_tables_%s_fast.cpp" % m.suffix: emit_tables_cpp(pf, "mistra_oc_tables_%s" % m.suffix),
            "onchip_tables_%s_strict.cpp" % m.suffix: emit_tables_cpp(ps, "mistra_oc_tables_%s" % m.suffix),
        }
        for fn, text in files.items():
            path = os.path.join(outdir, fn)
            if not (os.path.exists(path) and open(path).read() == text):
                with open(path, "w") as f:
                    f.write(text)
            print("wrote", path, len(text) // 1024, "KiB")


TAIL = {"gas": 32, "aer": 64}

if __name__ == "__main__":
    main(sys.argv)
