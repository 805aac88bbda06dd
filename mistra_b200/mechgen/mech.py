"""Mechanism tables (loaded from ``mistra_b200/mech/*.json``) plus the static
analyses the generators need.  Pure Python/numpy; no reference access."""
from __future__ import annotations

import functools
import json
import os

import numpy as np

MECH_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mech")
MECH_NAMES = ("gas", "aer", "tot")
MECH_ID = {"gas": 0, "aer": 1, "tot": 2}


class Mechanism:
    def __init__(self, d):
        self.d = d
        self.name = d["name"]
        self.suffix = d["suffix"]
        self.nvar = d["nvar"]
        self.nfix = d["nfix"]
        self.nreact = d["nreact"]
        self.nspec = d["nspec"]
        self.lu_nonzero = d["lu_nonzero"]
        self.reactions = d["reactions"]          # [[kind, idx|lit], ...] per reaction
        self.vdot = d["vdot"]                    # [[sign, coef|None, a_idx], ...] per species
        self.B = d["B"]                          # [[b_idx, factors], ...]
        self.bdim = d["bdim"]
        self.jvs = d["jvs"]                      # [[sign, coef|None, b_idx], ...] per nz
        self.icol = np.asarray(d["lu_icol"], dtype=np.int32)
        self.crow = np.asarray(d["lu_crow"], dtype=np.int32)
        self.diag = np.asarray(d["lu_diag"], dtype=np.int32)
        self.spc_names = d["spc_names"]
        self.eqn_names = d["eqn_names"]
        self.source = d["source"]

    # -- literal coefficients -------------------------------------------------
    @functools.cached_property
    def coef_literals(self):
        """Sorted list of distinct literal coefficients appearing in Vdot/JVS/B
        (strings exactly as written in the reference)."""
        s = set()
        for v in self.vdot:
            for _, c, _ in v:
                if c is not None:
                    s.add(c)
        for v in self.jvs:
            for _, c, _ in v:
                if c is not None:
                    s.add(c)
        for _, facs in self.B:
            for f in facs:
                if f[0] == "N":
                    s.add(f[1])
        for facs in self.reactions:
            for f in facs:
                if f[0] == "N":
                    s.add(f[1])
        return sorted(s, key=lambda x: (float(x), x))

    @staticmethod
    def literal_value(lit, f32):
        """Value of a Fortran default-kind literal.  Integer literals are exact;
        a REAL literal without a D exponent is binary32 under the reference's
        preferred flags (SURVEY §8a trap 1) and binary64 under -r8."""
        if lit.isdigit():
            return float(int(lit))
        if "d" in lit.lower():
            return float(lit.lower().replace("d", "e"))
        return float(np.float32(lit)) if f32 else float(lit)

    # -- LU structure ---------------------------------------------------------
    @functools.cached_property
    def row_of(self):
        r = np.zeros(self.lu_nonzero, dtype=np.int32)
        for k in range(self.nvar):
            r[self.crow[k]:self.crow[k + 1]] = k
        return r

    @functools.cached_property
    def pos(self):
        """dict (row, col) -> storage index"""
        return {(int(self.row_of[i]), int(self.icol[i])): i for i in range(self.lu_nonzero)}

    def decomp_ops(self):
        """Yield the elimination program of KppDecomp in reference order:
        for row k, for each strictly-lower entry kk (column j):
            ('piv', k, kk, diag_j)            a = W[j]/JVS[diag j]
            ('upd', k, dst_idx, kk, src_idx)  JVS[dst] -= L[k,j]*JVS[src]
        with dst_idx the storage slot of (k, col(src)) in row k."""
        pos = self.pos
        for k in range(self.nvar):
            for kk in range(self.crow[k], self.diag[k]):
                j = int(self.icol[kk])
                yield ("piv", k, kk, int(self.diag[j]))
                for jj in range(self.diag[j] + 1, self.crow[j + 1]):
                    c = int(self.icol[jj])
                    yield ("upd", k, pos[(k, c)], kk, jj)

    @functools.cached_property
    def decomp_counts(self):
        npiv = nupd = 0
        for op in self.decomp_ops():
            if op[0] == "piv":
                npiv += 1
            else:
                nupd += 1
        return npiv, nupd

    # -- algorithmic flop counts (SURVEY §8d convention: FMA = 2, div = 1) -----
    @functools.cached_property
    def flops(self):
        fun_mul = sum(len(r) - 1 for r in self.reactions)
        fun_agg_mul = sum(1 for v in self.vdot for _, c, _ in v if c is not None)
        fun_add = sum(max(len(v) - 1, 0) for v in self.vdot)
        fun = fun_mul + fun_agg_mul + fun_add
        jac_mul = sum(len(f) - 1 for _, f in self.B)
        jac_agg_mul = sum(1 for v in self.jvs for _, c, _ in v if c is not None)
        jac_add = sum(max(len(v) - 1, 0) for v in self.jvs)
        jac = jac_mul + jac_agg_mul + jac_add
        npiv, nupd = self.decomp_counts
        nlow = int(sum(self.diag[:self.nvar] - self.crow[:self.nvar]))
        nup = int(sum(self.crow[1:] - self.diag[:self.nvar] - 1))
        return {
            "fun": fun, "jac": jac,
            "prep": self.lu_nonzero + self.nvar,
            "decomp": npiv + 2 * nupd,
            "solve": 2 * (nlow + nup) + self.nvar,
            "vec": 30 * self.nvar,
            "decomp_div": npiv, "decomp_fma": nupd,
            "solve_fma": nlow + nup,
        }

    def flops_from_stats(self, stats):
        """Algorithmic flops of the reference formulation from the integrator's
        own counters (Nfun,Njac,Nstp,Nacc,Nrej,Ndec,Nsol,Nsng), SURVEY §8d."""
        f = self.flops
        stats = np.asarray(stats, dtype=np.float64).reshape(-1, 8)
        nfun, njac, nstp, ndec, nsol = (stats[:, 0].sum(), stats[:, 1].sum(), stats[:, 2].sum(),
                                        stats[:, 5].sum(), stats[:, 6].sum())
        return (nfun * f["fun"] + njac * f["jac"] + ndec * (f["prep"] + f["decomp"])
                + nsol * f["solve"] + nstp * f["vec"])

    # -- compulsory bytes per cell-integration --------------------
    @property
    def io_bytes(self):
        return 8 * (2 * self.nvar + self.nfix + self.nreact)


@functools.lru_cache(maxsize=None)
def load(name):
    with open(os.path.join(MECH_DIR, name + ".json")) as f:
        return Mechanism(json.load(f))
