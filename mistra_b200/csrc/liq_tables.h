// Per-layer tables of the liq_parm chain (row N2): henry_x, v_mean_x, st_coeff_x, equil_co_x (kpp.f90:664-2145,
// 2954-3363) as host/device functions generated from the reference's statements (mechgen/liqgen.py).  The generated
// statements keep the Fortran array notation; the macros below map it onto ONE layer's rows.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LIQ_HD __host__ __device__ __forceinline__
#else
#define LIQ_HD inline
#endif

struct LiqLayer {
  double tk;              // t(k) = tt(k): temperature of the layer [K]
  const double *conv2;    // [nkc]      COMMON /blck13/ conv2(kc,k) = 1/(1000 cw)
  const double *xgamma;   // [nkc][j6]  COMMON /kpp_mol/ xgamma(j,kc,k) activity coefficients
  double an2o5[4];        // a_n2o5(k,kc), kpp.f90:8377 (only read with lpJoyce14bc)
  int lpjoyce14bc, lpbuxmann15alph, j6, f32;   // config switches (USE config, ONLY : lpJoyce14bc, lpBuxmann15alph)
  double *henry, *vmean, *alpha;   // [NSPEC] rows of the layer
  double *xkef, *xkeb;             // [nkc][NSPEC]
};

// a_n2o5 (kpp.f90:8377-8439): N2O5 accommodation coefficient from the bin's nitrate / chloride molality and the water
// of bin 1; cw, cm [nkc], sion1_13 / sion1_14 = sion1(13,kc,k), sion1(14,kc,k)
LIQ_HD double liq_a_n2o5(const double *cw, const double *cm, int kc, double sion1_13, double sion1_14)
{
  double xno3m = 0.0, xclm = 0.0, xh2o = 0.0, denom = 1.0;
  if (cw[kc] > 0.0) {
    xno3m = sion1_13 / cw[kc] * 1e-3;
    xclm = sion1_14 / cw[kc] * 1e-3;
  }
  if (cm[0] > 0.0 && cw[0] > 0.0) xh2o = 55.55 * (cm[0] / cw[0]);
  const double xk2f = 1.15e6 - 1.15e6 * exp(-0.13 * xh2o);
  if (xno3m > 0.0) denom = 1.0 + 6.e-2 * xh2o / xno3m + 29.0 * xclm / xno3m;
  return 3.2e-8 * xk2f * (1.0 - (1.0 / denom));
}

#define RL(x) (L_.f32 ? (double)x##f : (double)x)
#define henry(i, k) L_.henry[(i)-1]
#define vmean(i, k) L_.vmean[(i)-1]
#define alpha(i, k) L_.alpha[(i)-1]
#define xkef(i, kc, k) L_.xkef[((kc)-1) * nspec + (i)-1]
#define xkeb(i, kc, k) L_.xkeb[((kc)-1) * nspec + (i)-1]
#define tt(k) L_.tk
#define t(k) L_.tk
#define conv2(kc, k) L_.conv2[(kc)-1]
#define xgamma(j, kc, k) L_.xgamma[((kc)-1) * L_.j6 + (j)-1]
#define a_n2o5(k, kc) L_.an2o5[(kc)-1]
#include "_gen/liq_tables.inc"
#undef RL
#undef henry
#undef vmean
#undef alpha
#undef xkef
#undef xkeb
#undef tt
#undef t
#undef conv2
#undef xgamma
#undef a_n2o5

// all four routines for one layer; mech 1 = aer (2 bins), 2 = tot (4 bins).  Outputs are overwritten where the
// reference assigns them; xkef / xkeb of a bin without liquid water (conv2 <= 0) are left as they are.
LIQ_HD void liq_tables_layer(int mech, LiqLayer &L)
{
  if (mech == 1) { liq_henry_a(L); liq_v_mean_a(L); liq_st_coeff_a(L); liq_equil_co_a(L); }
  else { liq_henry_t(L); liq_v_mean_t(L); liq_st_coeff_t(L); liq_equil_co_t(L); }
}
