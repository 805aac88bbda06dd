/* mistra_liq.h - the per-layer tables of the liq_parm chain (part of libmistra_kpp.so; host version in
 * libmistra_rconst.so).
 *
 * "Next" row N2 of the hot-path scope, the part without reductions over the spectrum: for every layer
 *   henry_a / henry_t        /root/reference/src/kpp.f90:1914 / 1676   henry(NSPEC)   (inverse Henry constants, dimensionless)
 *   v_mean_a / v_mean_t      kpp.f90:1472 / 1268                        vmean(NSPEC)   mean molecular speed [m/s]
 *   st_coeff_a / st_coeff_t  kpp.f90:857 / 664                          alpha(NSPEC)   accommodation coefficients
 *   equil_co_a / equil_co_t  kpp.f90:3162 / 2954                        xkef, xkeb(NSPEC,nkc)  equilibrium rates
 * from the layer's temperature, conv2(nkc) = 1/(1000 cw) (COMMON /blck13/), the activity coefficients xgamma(j6,nkc)
 * (COMMON /kpp_mol/) and, with the configuration switch lpJoyce14bc, cw / cm / sion1(13:14) for a_n2o5
 * (kpp.f90:8377).  The formulas are generated from the reference's statements (mechgen/liqgen.py).  Together with
 * mistra_cwrc (cw, cm, conv2) and mistra_fastkmt (xkmt) this is everything liq_parm (kpp.f90:516-657) hands to
 * Update_RCONST_a/_t, so alpha, vmean, henry, xkef, xkeb no longer cross PCIe.
 *
 * Rows are per cell (= layer of a column).  mech: 1 = aer (nkc = 2 bins), 2 = tot (nkc = 4).  Outputs [ncell][NSPEC] and
 * [ncell][nkc][NSPEC] - the layout of include/mistra_rconst.h - are zero-filled first, then assigned where the reference
 * assigns; the reference's loop bounds (st_coeff, equil_co start at layer 2, bins 1..2 only in equil_co_a) are the
 * caller's choice of cells.  Returns 0 or MISTRA_KPP_E*. */
#ifndef MISTRA_LIQ_H
#define MISTRA_LIQ_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mistra_liq_args {
  const double *t;       /* [ncell]            */
  const double *conv2;   /* [ncell][nkc]       */
  const double *xgamma;  /* [ncell][nkc][j6]   */
  const double *cw;      /* [ncell][nkc]  only read with lpjoyce14bc (may be NULL otherwise) */
  const double *cm;      /* [ncell][nkc]  "                                                   */
  const double *sion1_13_14; /* [ncell][nkc][2] sion1(13,kc,k), sion1(14,kc,k)  "             */
  int32_t j6;            /* 55 */
  int32_t lpjoyce14bc;   /* config switch (USE config, ONLY : lpJoyce14bc) */
  int32_t f32_literals;  /* 1: default-REAL literals binary32 (reference's preferred flags) */
  int32_t lpbuxmann15alph; /* config switch (st_coeff_t) */
  double *henry, *vmean, *alpha; /* [ncell][NSPEC] out */
  double *xkef, *xkeb;           /* [ncell][nkc][NSPEC] out */
} mistra_liq_args;

/* device arrays on the current device, asynchronous on `stream` */
int mistra_liq_tables_device(int mech, int64_t ncell, const mistra_liq_args *a, void *stream);
/* host arrays, plain C++ (libmistra_rconst.so); nthreads <= 1: serial */
int mistra_liq_tables_host(int mech, int64_t ncell, const mistra_liq_args *a, int nthreads);
int64_t mistra_liq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
