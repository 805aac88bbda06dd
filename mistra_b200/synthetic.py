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

    def compact_rates(self, alloc=None, idx=None):
        """The inputs of Update_RCONST_g in the compact form of include/mistra_kpp_rates.h."""
        from . import kpp
        return kpp.CompactRates(0, self.cb1, self.scal, self.ph_rat, ycwd=self.ycwd, f32_literals=self.f32, alloc=alloc,
                                idx=idx, yhenry=self.yhenry, yxkmtd=self.yxkmtd, yxeq=self.yxeq)

    def rconst(self, var=None, sl=slice(None)):
        """Update_RCONST_g for the cells in `sl` at concentrations `var`."""
        c = self.conc(var)[sl]
        return rc.update_rconst(0, self.cb1[sl], self.scal[sl], self.ph_rat[sl], c,
                                yhenry=self.yhenry[sl], yxkmtd=self.yxkmtd[sl], yxeq=self.yxeq[sl],
                                ycwd=self.ycwd[sl], f32_literals=self.f32)


# ---------------------------------------------------------------------------------
# Aqueous-phase inputs (aer mechanism).  The reference produces them in its liq_parm
# chain (kpp.f90:516-657): henry_a 1914, v_mean_a 1472, st_coeff_a 857, fast_k_mt_a
# 2683, equil_co_a 3162, cw_rc 2152.  That chain is "next" row N2; until
# it lands the generator restates its tables for a mono-disperse aerosol per bin
# (effective radius r, liquid water content cw) with activity coefficients = 1.
# Tables below: (k_H^cp at 298 K in M/atm, -dlnk/d(1/T)), molar mass in kg/mol,
# accommodation coefficient (temperature dependence dropped where it is weak).
HENRY = {"NO": (1.9e-3, 1480.), "NO2": (6.4e-3, 2500.), "HNO3": (2.5e6 / 1.5e1, 8694.),
         "HNO4": (1.2e4, 6900.), "NH3": (58., 4085.), "SO2": (1.2, 3120.), "O3": (1.2e-2, 2560.),
         "ACO2": (3.7e3, 5700.), "ACTA": (4.1e3, 6300.), "HCHO": (7.0e3, 6425.), "ALD2": (1.3e1, 5700.),
         "H2O2": (1.0e5, 6338.), "ROOH": (3.0e2, 5322.), "HONO": (4.9e1, 4780.), "PAN": (2.8, 6500.),
         "HCl": (2. / 1.7, 9001.), "NO3": (2., 2000.), "DMS": (4.8e-1, 3100.), "DMSO": (5.e4, 6425.),
         "HOCl": (6.7e2, 5862.), "Cl2": (9.1e-2, 2500.), "HBr": (1.3, 10239.), "Br2": (7.6e-1, 4094.),
         "BrCl": (9.4e-1, 5600.), "HOBr": (9.3e1, 5862.), "I2": (3., 4431.), "HOI": (4.5e2, 5862.),
         "ICl": (1.1e2, 5600.), "IBr": (2.4e1, 5600.), "CH3I": (1.4e-1, 4300.), "CH2I2": (2.3, 5000.),
         "CH2ClI": (8.9e-1, 4300.), "OH": (3.0e1, 4300.), "HO2": (3.9e3, 5900.), "MO2": (6., 5600.),
         "IO": (4.5e2, 5862.), "CO2": (3.1e-2, 2423.), "CO": (9.9e-4, 1300.), "O2": (1.3e-3, 1500.),
         "CH3OH": (1.6e2, 5600.), "C2H5OH": (1.5e2, 6400.), "H2": (7.8e-4, 500.),
         "XOR": (1.5e2, 6400.), "SOR": (1.5e2, 6400.)}
HENRY_CONST = {"H2SO4": 1.e16, "CH4": 1.3e-3, "C2H6": 2.0e-3, "ETHE": 4.9e-3, "C3H7I": 1.1e-1,
               "DMSO2": 1.e16, "CH3SO2H": 1.e16, "CH3SO3H": 1.e16, "ClONO": 4.6e-2}
MOLMASS = {"NO": 3.e-2, "NO2": 4.6e-2, "HNO3": 6.3e-2, "NH3": 1.7e-2, "SO2": 6.4e-2, "H2SO4": 9.8e-2,
           "O3": 4.8e-2, "ACO2": 4.6e-2, "ACTA": 6.e-2, "HCHO": 3.e-2, "H2O2": 3.4e-2, "ROOH": 4.8e-2,
           "HONO": 4.7e-2, "HCl": 3.6e-2, "N2O5": 1.08e-1, "HNO4": 7.9e-2, "NO3": 6.2e-2, "DMS": 6.2e-2,
           "HOCl": 5.2e-2, "ClNO3": 9.7e-2, "Cl2": 7.1e-2, "HBr": 8.1e-2, "HOBr": 9.7e-2,
           "BrNO3": 1.42e-1, "Br2": 1.6e-1, "BrCl": 1.15e-1, "HI": 1.28e-1, "HOI": 1.44e-1,
           "I2O2": 2.86e-1, "INO2": 1.73e-1, "INO3": 1.89e-1, "I2": 2.54e-1, "ICl": 1.62e-1,
           "IBr": 2.07e-1, "HIO3": 1.76e-1, "DMSO": 7.8e-2, "DMSO2": 9.4e-2, "CH3SO2H": 8.0e-2,
           "CH3SO3H": 9.6e-2, "CO2": 4.4e-2, "CH3OH": 3.2e-2, "C2H5OH": 4.6e-2, "XOR": 1.09e-1,
           "SOR": 9.4e-2, "OH": 1.7e-2, "HO2": 3.3e-2, "MO2": 4.7e-2, "O2": 3.2e-2, "IO": 1.43e-1,
           "OIO": 1.59e-1}
ALPHA = {"H2SO4": 0.65, "O3": 2.0e-3, "O2": 1.0e-2, "OH": 1.0e-2, "HO2": 2.0e-1, "H2O2": 0.11,
         "NO": 5.0e-5, "NO2": 1.5e-3, "NO3": 4.0e-2, "N2O5": 0.1, "HONO": 4.0e-2, "HNO3": 0.5,
         "NH3": 6.0e-2, "MO2": 1.0e-2, "ROOH": 5.0e-3, "HCHO": 4.0e-2, "ACO2": 1.4e-2, "ACTA": 6.7e-2,
         "CH3OH": 5.6e-2, "C2H5OH": 4.8e-2, "CO2": 1.0e-2, "HCl": 0.1, "Cl2": 3.8e-2, "HBr": 3.0e-2,
         "HOBr": 0.6, "HOCl": 0.6, "BrNO3": 0.8, "ClNO3": 0.1, "Br2": 3.8e-2, "BrCl": 0.33,
         "SO2": 0.11, "CH3SO3H": 7.6e-2, "DMS": 1.0e-2, "DMSO": 4.8e-2, "DMSO2": 3.0e-2,
         "CH3SO2H": 2.0e-4, "INO3": 0.1, "HOI": 0.6, "HI": 3.6e-2, "I2": 1.0e-2, "IO": 0.5,
         "I2O2": 0.1, "ICl": 1.0e-2, "IBr": 1.0e-2, "INO2": 0.1, "OIO": 1.0, "HIO3": 1.0e-2,
         "XOR": 7.0e-2}
# species exchanged between gas and aqueous phase (lex of fast_k_mt_a, kpp.f90:2683)
LEX = ("NO2", "HNO3", "NH3", "SO2", "H2SO4", "O3", "ACO2", "HCHO", "H2O2", "HONO", "HCl", "N2O5",
       "HNO4", "NO3", "OH", "HO2", "MO2", "CO2", "O2", "ROOH", "HOCl", "Cl2", "HBr", "HOBr", "Br2",
       "BrCl", "DMSO", "ClNO3", "BrNO3", "CH3SO3H", "DMS", "CH3SO2H", "DMSO2", "HOI", "IO", "I2",
       "ICl", "IBr", "OIO", "INO2", "INO3", "HI", "I2O2", "HIO3", "NO", "ACTA", "CH3OH", "C2H5OH",
       "XOR", "SOR")
# acid-base / halogen equilibria: species -> (kef0, kef T-coefficient or None, keb0, keb T-coeff or
# None, keb multiplied by cv2?, kef multiplied by cv2?)   (equil_co_a, kpp.f90:3162-3368)
EQUIL = {"H2O": (1.0e-5, -6716., 1.0e9, None, True, False), "HO2": (1.6e5, None, 1.e10, None, True, False),
         "ACO2": (1.8, None, 1.0e4, None, True, False), "CO2": (4.3e-2, -913., 1.0e5, None, True, False),
         "HONO": (5.1e3, -1260., 1.0e7, None, True, False), "HNO3": (1.54e10, 8700., 1.0e9, None, True, False),
         "HNO4": (2.0e3, None, 2.0e8, None, True, False), "NH3": (1.7e5, -4325., 1.0e10, None, True, False),
         "HSO3ml%d": (6.0e2, 1120., 1.0e10, None, True, False), "H2SO4": (1.0e12, None, 1.0e9, None, True, False),
         "HSO4ml%d": (1.02e6, 2720., 1.0e8, None, True, False), "SO2": (1.7e8, 2090., 1.0e10, None, True, False),
         "HCHO": (1.e10, None, 1.e5, None, False, True), "HCl": (1.7e10, 6896., 1.0e4, None, True, False),
         "Cl2ml%d": (5.2e4, None, 1.e10, None, True, False), "HOCl": (3.2e2, None, 1.0e10, None, True, False),
         "HBr": (1.0e13, None, 1.0e4, None, True, False), "Br2": (2.95e4, -4068., 1.17e10, -1812., True, False),
         "HOBr": (2.3e1, -3091., 1.0e10, None, True, False),
         "BrCl2ml%d": (5.e9, 1143., 1.3e9, None, False, True), "Br2Clml%d": (5.e9, None, 2.8e5, None, False, True),
         "Br2l%d": (5.e9, None, 3.85e9, None, False, True), "ICl": (1.0e11, None, 1.3e9, None, False, True),
         "IBr": (1.0e11, None, 3.5e8, None, False, True), "IClBrml%d": (5.e9, None, 2.8e5, None, False, True),
         "I2": (5.e9, None, 3.85e9, None, False, True), "HIO3": (1.57e4, None, 1.0e5, None, True, False)}


class _AqueousEnsemble:
    """Per-cell inputs of a mechanism with aqueous chemistry for the layers below nf of
    `ncol` columns: 98 cells per column (k = 2..99; kpp.f90:4381-4391 - layers k >= nf
    are always gas-only).  Bins 1/2 = deliquesced sulfate / sea-salt aerosol, bins 3/4
    (tot only) = the droplets grown on them."""

    LAYERS = NF - 2
    MECH = None      # (id, name, number of aqueous bins)

    def __init__(self, ncol, seed=SEED, halo=True, iod=True, f32_literals=1, col0=0):
        g = GasEnsemble(ncol, seed=seed, halo=halo, iod=iod, f32_literals=f32_literals, col0=col0)
        self.mech, mname, nb = self.MECH
        self.nbins = nb
        m = self.m = mechmod.load(mname)
        self.f32 = f32_literals
        nspec = m.nvar + m.nfix
        idx = self.idx = {n: i for i, n in enumerate(m.spc_names)}
        L = self.LAYERS
        sel = (np.arange(ncol)[:, None] * CELLS_PER_COLUMN + np.arange(L)[None, :]).reshape(-1)
        ncell = self.ncell = ncol * L
        self.ncol = ncol
        self.cb1 = g.cb1[sel]
        self.ph_rat = g.ph_rat[sel]
        tt, pk = self.cb1[:, 1], self.cb1[:, 3]
        f32 = (lambda x: float(np.float32(x))) if f32_literals else float
        # liquid water and effective radius per bin
        aq = np.empty((ncol, 12))
        for c in range(ncol):
            aq[c] = np.random.default_rng([seed, col0 + c, 1]).uniform(0.0, 1.0, 12)
        aq = np.repeat(aq, L, axis=0)
        cw = np.stack([10.0 ** (-11.5 + aq[:, 0]), 10.0 ** (-10.5 + aq[:, 1]),
                       10.0 ** (-8.3 + aq[:, 6]), 10.0 ** (-8.3 + aq[:, 7])], axis=1)[:, :nb]     # m3/m3
        rr = np.stack([8.0e-8 * 5.0 ** aq[:, 2], 8.0e-7 * 5.0 ** aq[:, 3],
                       3.0e-6 * 3.0 ** aq[:, 8], 5.0e-6 * 3.0 ** aq[:, 9]], axis=1)[:, :nb]       # m
        cv2 = 1.0e-3 / cw                                                                # conv2
        self.ycw = cw
        scal = np.zeros((ncell, 13))
        scal[:, 0] = CONV1
        scal[:, 1] = 1.0 if halo else 0.0
        scal[:, 2] = 1.0 if iod else 0.0
        scal[:, 5:5 + nb] = 1.0                 # xliq = 1 -> xhet1 = xhet2 = 0
        scal[:, 9:9 + nb] = cv2
        self.scal = scal
        # inverse dimensionless Henry constants (henry_a / henry_t)
        tfact = 1.0 / tt - 3.3540e-3
        fct = 0.0820577 * tt
        yhenry = np.zeros((ncell, nspec))
        for s, (a0, b0) in HENRY.items():
            if s in idx:
                yhenry[:, idx[s]] = 1.0 / (a0 * np.exp(b0 * tfact) * fct)
        for s, a0 in HENRY_CONST.items():
            if s in idx:
                yhenry[:, idx[s]] = 1.0 / (a0 * fct)
        # mass-transfer coefficients (fast_k_mt_a / _t for one radius per bin)
        freep = 2.28e-5 * tt / pk
        yxkmt = np.zeros((ncell, nb, nspec))
        for s in LEX:
            if s not in idx:
                continue
            vmean = np.sqrt(tt / MOLMASS.get(s, 5.0e-2)) * 4.60138
            x1 = 4.0 / (3.0 * ALPHA.get(s, 0.1))
            for kc in range(nb):
                yxkmt[:, kc, idx[s]] = vmean / (rr[:, kc] * (rr[:, kc] / freep + x1))
        # equilibrium rate coefficients (equil_co_a / _t, activity coefficients = 1)
        ykef = np.zeros((ncell, nb, nspec))
        ykeb = np.zeros((ncell, nb, nspec))
        tf2 = 1.0 / tt - 3.354e-3
        for s, (f0, fb, b0, bb, bcv, fcv) in EQUIL.items():
            name = s % (1,) if "%d" in s else s          # the reference indexes by the bin-1 name
            if name not in idx:
                continue
            for kc in range(nb):
                kf = f0 * (np.exp(fb * tf2) if fb is not None else 1.0)
                kb = b0 * (np.exp(bb * tf2) if bb is not None else 1.0)
                ykef[:, kc, idx[name]] = kf * (cv2[:, kc] if fcv else 1.0)
                ykeb[:, kc, idx[name]] = kb * (cv2[:, kc] if bcv else 1.0)
        self.yhenry, self.yxkmt, self.ykef, self.ykeb = yhenry, yxkmt, ykef, ykeb
        # dry-aerosol arrays are multiplied by xhet = 0 here; keep the gas ensemble's values
        gi = g.idx
        self.yxkmtd = np.zeros((ncell, 2, nspec))
        self.yxeq = np.zeros((ncell, nspec))
        for s in ("HNO3", "N2O5", "NH3", "H2SO4"):
            self.yxkmtd[:, :, idx[s]] = g.yxkmtd[sel][:, :, gi[s]]
        self.yxeq[:, idx["HNO3"]] = g.yxeq[sel][:, gi["HNO3"]]
        self.ycwd = g.ycwd[sel]
        # FIX = O2, H2O, N2, H2Ol1.. with FIX(H2Olz) = 55.55/cvvz (aer.f:197-206)
        assert m.spc_names[m.nvar:] == ["O2", "H2O", "N2"] + ["H2Ol%d" % (k + 1) for k in range(nb)]
        gf = g.fix[sel]
        self.fix = np.concatenate([gf, f32(55.55) / cv2], axis=1)
        # VAR: gas phase from the gas ensemble (by name), ions per initc (kpp.f90:337-381)
        var = np.zeros((ncell, m.nvar))
        aqs = tuple("l%d" % (k + 1) for k in range(4))
        for s, i in gi.items():
            if i < g.m.nvar and s in idx and idx[s] < m.nvar and not s.endswith(aqs):
                var[:, idx[s]] = g.var[sel][:, i]
        # aerosol: 2..6 mol/l of salt; droplets: the same kind of salt diluted to 1e-4..1e-3 mol/l
        molar = np.stack([2.0 + 4.0 * aq[:, 4], 2.0 + 4.0 * aq[:, 5],
                          10.0 ** (-4.0 + aq[:, 10]), 10.0 ** (-4.0 + aq[:, 11])], axis=1)[:, :nb]
        x0 = molar * 1.0e3 * cw                              # mol/m3(air)
        xi = 1.0 if iod else 0.0
        xso4, xhco3, xno3, xbr = 0.0485, 4.2e-3, 1.0e-7, 1.45e-3
        xim, xio3 = 7.4e-8 / 0.545 * xi, 2.64e-7 / 0.545 * xi
        xcl = 1.0 - (xso4 + xhco3 + xno3 + xbr + xim + xio3)
        sulf = (("NH4pl", 1.34), ("SO42ml", 0.34), ("NO3ml", 0.004), ("HSO4ml", 0.656))
        salt = (("SO42ml", xso4), ("HCO3ml", xhco3), ("NO3ml", xno3), ("Clml", xcl), ("Brml", xbr),
                ("Iml", xim), ("IO3ml", xio3), ("DOMl", 0.27 * xbr))
        for kc in range(nb):
            for s, f in (sulf if kc % 2 == 0 else salt):
                var[:, idx["%s%d" % (s, kc + 1)]] = f * x0[:, kc]
        self.var = var

    def conc(self, var=None):
        return np.concatenate([self.var if var is None else var, self.fix], axis=1)

    def compact_rates(self, alloc=None, idx=None):
        """The inputs of Update_RCONST_a / _t in the compact form of include/mistra_kpp_rates.h."""
        from . import kpp
        return kpp.CompactRates(self.mech, self.cb1, self.scal, self.ph_rat, ycw=self.ycw, ycwd=self.ycwd,
                                f32_literals=self.f32, alloc=alloc, idx=idx, yhenry=self.yhenry, yxkmt=self.yxkmt, ykef=self.ykef,
                                ykeb=self.ykeb, yxkmtd=self.yxkmtd, yxeq=self.yxeq)

    def rconst(self, var=None, sl=slice(None)):
        """Update_RCONST_a / _t for the cells in `sl` at concentrations `var`."""
        c = self.conc(var)[sl]
        return rc.update_rconst(self.mech, self.cb1[sl], self.scal[sl], self.ph_rat[sl], c,
                                yhenry=self.yhenry[sl], yxkmt=self.yxkmt[sl], ykef=self.ykef[sl],
                                ykeb=self.ykeb[sl], yxkmtd=self.yxkmtd[sl], yxeq=self.yxeq[sl],
                                ycw=self.ycw[sl], ycwd=self.ycwd[sl], f32_literals=self.f32)


class AerEnsemble(_AqueousEnsemble):
    """aer mechanism: gas + aqueous chemistry in the aerosol bins 1, 2."""
    MECH = (1, "aer", 2)


class TotEnsemble(_AqueousEnsemble):
    """tot mechanism: gas + aerosol bins 1, 2 + droplet bins 3, 4 (cloudy layers)."""
    MECH = (2, "tot", 4)
