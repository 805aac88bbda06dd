"""Synthetic ensemble of independent Mistra columns.

Every column has the reference's 150-level vertical grid (148 chemistry cells,
layers k = 2..149, kpp.f90:4294-4310).  Per-cell inputs are produced the way the
reference's kpp_driver does (kpp.f90:4315-4468): temperature / air density /
water vapour, photolysis frequencies (zero for night columns), halogen switches,
dry-aerosol uptake coefficients (dry_rates_g, kpp.f90:4697-4858) and initial gas
concentrations interpolated like initc (kpp.f90:229-279) with log-normal jitter.

Only numpy + the host rate-constant library are used here; no reference files
are read at run time.
"""
from __future__ import annotations

import numpy as np

from . import rconst as rc
from .mechgen import mech as mechmod

SEED = 20261018
N_LEVELS = 150          # global_params.f90: n
NF = 100                # global_params.f90: nf
CELLS_PER_COLUMN = 148  # k = 2 .. n-1
AVOGADRO = 6.022140857e+23
CONV1 = AVOGADRO * 1.0e-6
M_AIR = 28.96546e-3
GAS_CONST = 8.3144743

# src/mech/gas_species.csv: non-zero initial mixing ratios (ground, top) in ppb
GAS_PPB = {"NO2": (0.02, 0.03), "HNO3": (0.01, 0.05), "NH3": (0.08, 0.08), "SO2": (0.09, 0.045),
           "O3": (55.0, 70.0), "CH4": (1800.0, 1800.0), "C2H6": (0.5, 0.5), "HCHO": (0.1, 0.1),
           "H2O2": (0.4, 0.2), "PAN": (0.01, 0.1), "HCl": (0.005, 0.005), "Br2": (0.001, 0.001),
           "CH3I": (0.002, 0.002), "C3H7I": (0.001, 0.001), "CO": (130.0, 130.0),
           "CO2": (388000.0, 388000.0), "H2": (500.0, 500.0)}
HALOGEN_BL_ONLY = {"Br2", "CH3I", "C3H7I"}   # gas_is_halo and not HCl: no gradient, BL only

# clear-sky noon photolysis frequencies J0 [1/s], index = jrate.f:342-388 order (1-based there)
J0 = np.zeros(47)
for _i, _v in {1: 6.3e-3, 2: 2.0e-2, 3: 2.0e-5, 4: 1.3e-3, 5: 5.0e-7, 6: 5.5e-6, 7: 2.7e-6,
               8: 2.5e-5, 9: 3.5e-5, 10: 1.6e-1, 11: 1.3e-6, 12: 3.5e-5, 13: 2.2e-4, 14: 3.5e-5,
               15: 1.0e-3, 16: 1.2e-3, 17: 4.5e-6, 18: 3.0e-4, 19: 1.55e-3, 20: 1.7e-3,
               21: 7.0e-4, 22: 2.5e-2, 23: 8.0e-3, 24: 3.0e-2, 25: 1.5e-1, 26: 6.0e-3,
               27: 1.2e-1, 28: 2.0e-2, 29: 6.0e-2, 30: 3.0e-3, 31: 3.0e-6, 32: 1.0e-5,
               33: 9.0e-5, 34: 5.0e-3, 35: 7.0e-2, 37: 2.0e-3, 38: 2.7e-5, 39: 4.0e-7,
               47: 4.5e-4}.items():
    J0[_i - 1] = _v
J0[36 - 1] = 9.0 * J0[16 - 1]
J0[40 - 1] = J0[35 - 1]
J0[44 - 1] = J0[34 - 1] / 17.0
J0[46 - 1] = J0[31 - 1]


def vertical_grid(detamin=10.0, etaw1=2000.0):
    """Layer mid heights eta(k), k = 1..n (str.f90:1570-1621): equidistant below
    eta(nf), logarithmically stretched above up to etaw1."""
    n, nf = N_LEVELS, NF
    etw = np.zeros(n + 1)           # 1-based
    for k in range(2, nf + 1):
        etw[k] = (k - 1) * detamin
    x0, x1, x3 = detamin, etaw1, 1.0
    j = 0
    while x1 > etaw1 - etw[nf]:
        x0 += detamin
        j += 1
        x3 = detamin / x0 + 1.0
        etw[nf + 1] = x0
        for k in range(nf + 2, n + 1):
            etw[k] = etw[k - 1] * x3
        x1 = etw[n] - etw[nf + 1]
        if j > 10000:
            raise RuntimeError("vertical grid did not converge")
    x0 = nf * detamin - etw[nf + 1]
    etw[nf + 1:] += x0
    eta = np.zeros(n + 1)
    for k in range(2, n + 1):
        eta[k] = 0.5 * (etw[k] + etw[k - 1])
    return eta[1:]                  # eta[0] = level 1


class GasEnsemble:
    """Per-cell inputs of the gas mechanism for `ncol` columns (ncol*148 cells)."""

    def __init__(self, ncol, seed=SEED, halo=True, iod=True, f32_literals=1, col0=0):
        self.mech = 0
        self.m = mechmod.load("gas")
        self.f32 = f32_literals
        m = self.m
        nspec = m.nvar + m.nfix
        idx = {n: i for i, n in enumerate(m.spc_names)}
        self.idx = idx
        # one independent stream per column so that shards of a larger ensemble
        # (multi-GPU) draw exactly the columns they own
        eta = vertical_grid()[1:N_LEVELS - 1]          # layers k = 2..149
        ncell = ncol * CELLS_PER_COLUMN
        self.ncol, self.ncell = ncol, ncell
        T0 = np.empty(ncol); p0 = np.empty(ncol); rh = np.empty(ncol); mu0 = np.empty(ncol)
        jit = np.empty((ncol, len(GAS_PPB))); aer = np.empty((ncol, 10))
        for c in range(ncol):
            r = np.random.default_rng([seed, col0 + c])
            T0[c] = r.uniform(255.0, 295.0)
            p0[c] = r.uniform(98.0e3, 102.5e3)
            rh[c] = r.uniform(0.3, 0.99)
            mu0[c] = r.uniform(-0.2, 1.0)
            jit[c] = r.lognormal(0.0, 0.5, len(GAS_PPB))
            aer[c] = r.uniform(0.0, 1.0, 10)
        z = eta[None, :]
        te = T0[:, None] - 6.5e-3 * z                                  # lapse 6.5 K/km
        pk = p0[:, None] * (te / T0[:, None]) ** (9.80665 * M_AIR / (GAS_CONST * 6.5e-3))
        air = pk / (GAS_CONST * te)                                    # mol/m3
        aircc = air * CONV1                                            # molec/cm3
        # saturation vapour pressure (Magnus) -> h2o mol/m3
        es = 610.94 * np.exp(17.625 * (te - 273.15) / (te - 30.11))
        h2o = rh[:, None] * es / (GAS_CONST * te)
        h2oppm = h2o / air * 1.0e6
        self.cb1 = np.stack([aircc, te, h2oppm, pk], axis=-1).reshape(ncell, 4)
        xhal, xiod = (1.0 if halo else 0.0), (1.0 if iod else 0.0)
        scal = np.zeros((ncell, 13))
        scal[:, 0] = CONV1
        scal[:, 1] = xhal
        scal[:, 2] = xiod
        scal[:, 3] = 1.0    # gas path: xhet1 = xhet2 = 1 (kpp.f90:4466-4467)
        scal[:, 4] = 1.0
        self.scal = scal
        # photolysis: all J = 0 when the sun is too low (kpp.f90:4343-4359)
        u0min = 3.48e-2
        day = np.where(mu0 >= u0min, np.maximum(mu0, 0.0), 0.0)
        self.ph_rat = np.repeat(day[:, None] * J0[None, :], CELLS_PER_COLUMN, axis=0)
        self.is_day = np.repeat(day > 0.0, CELLS_PER_COLUMN)
        # dry aerosol (dry_rates_g, kpp.f90:4697-4858)
        lw = 10.0 ** (-12.0 + 2.0 * aer[:, 0:2])                       # cwd  m3/m3  logU(1e-12,1e-10)
        rcd = 5.0e-8 * (2.0e-6 / 5.0e-8) ** aer[:, 2:4]                # mean radius logU(5e-8,2e-6) m
        ycwd = np.repeat(lw, CELLS_PER_COLUMN, axis=0)
        rcd = np.repeat(rcd, CELLS_PER_COLUMN, axis=0)
        tt = self.cb1[:, 1]
        freep = 2.28e-5 * tt / self.cb1[:, 3]
        f32 = (lambda x: float(np.float32(x))) if f32_literals else float
        xeq_hno3 = 1.54e+1 * np.exp(8700.0 * (1.0 / tt - 3.354e-3))
        h = (2.5e6 / xeq_hno3) * np.exp(8694.0 * (1.0 / tt - 3.3557e-3))
        henry_hno3 = 1.0 / (h * (f32(0.0820577) * tt))
        yxkmtd = np.zeros((ncell, 2, nspec))
        yhenry = np.zeros((ncell, nspec))
        yxeq = np.zeros((ncell, nspec))
        yhenry[:, idx["HNO3"]] = henry_hno3
        yxeq[:, idx["HNO3"]] = xeq_hno3
        for spc, mw, gam in (("HNO3", 6.3e-2, 0.02), ("N2O5", 1.08e-1, 0.02), ("NH3", 1.7e-2, 0.05),
                             ("H2SO4", 9.8e-2, 0.1)):
            vmean = np.sqrt(tt / mw) * f32(4.60138)
            for kc in range(2):
                r = rcd[:, kc]
                yxkmtd[:, kc, idx[spc]] = vmean * (1.0 / (r * (r / freep + 4.0 / (3.0 * gam))))
        self.yxkmtd, self.yhenry, self.yxeq, self.ycwd = yxkmtd, yhenry, yxeq, ycwd
        # FIX = (O2, H2O, N2) with the reference's default-REAL 0.21/0.79 (gas.f:146-148)
        airf = air.reshape(ncell)
        self.fix = np.stack([f32(0.21) * airf, h2o.reshape(ncell), f32(0.79) * airf], axis=-1)
        assert m.spc_names[m.nvar:] == ["O2", "H2O", "N2"]
        # initial VAR (initc, kpp.f90:229-279): ppb -> mol/m3 with exponential profile
        var = np.zeros((ncell, m.nvar))
        x4 = np.minimum(1.0, eta / 1900.0)
        xm = (air * 1.0e-9)                                             # [ncol,148]
        kinv_height = 700.0                                             # BL top for halogens
        for gi, (spc, (grd, top)) in enumerate(GAS_PPB.items()):
            x2 = -np.log(grd) + np.log(top + 1.0e-10)
            prof = grd * np.exp(x4 * x2)[None, :] * xm * jit[:, gi:gi + 1]
            if spc in HALOGEN_BL_ONLY:
                prof = np.where(eta[None, :] < kinv_height, grd * xm[:, :1] * jit[:, gi:gi + 1], 0.0)
            var[:, idx[spc]] = prof.reshape(ncell)
        # dry-aerosol nitrate / ammonium / sulfate seen by the gas mechanism (gas.f:151-156)
        for j, spc in enumerate(("HNO3l1", "NH3l1", "SO4l1", "HNO3l2", "NH3l2", "SO4l2")):
            v = 10.0 ** (-11.0 + 2.0 * aer[:, 4 + j])                   # mol/m3(air)
            var[:, idx[spc]] = np.repeat(v, CELLS_PER_COLUMN)
        self.var = var

    def conc(self, var=None):
        return np.concatenate([self.var if var is None else var, self.fix], axis=1)

    def rconst(self, var=None, sl=slice(None)):
        """Update_RCONST_g for the cells in `sl` at concentrations `var`."""
        c = self.conc(var)[sl]
        return rc.update_rconst(0, self.cb1[sl], self.scal[sl], self.ph_rat[sl], c,
                                yhenry=self.yhenry[sl], yxkmtd=self.yxkmtd[sl], yxeq=self.yxeq[sl],
                                ycwd=self.ycwd[sl], f32_literals=self.f32)
