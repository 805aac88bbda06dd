/* mistra_rconst.h - C ABI of the host-side rate-constant producer (libmistra_rconst.so).
 *
 * Restatement of Update_RCONST_g/_a/_t (/root/reference/src/gas.f:275-666,
 * aer.f:304-1400, tot.f:1040-2805) and of the rate-law functions they call
 * (/root/reference/src/kpp.f90:7127-8373), batched over cells.  In the reference
 * this runs on the host right before INTEGRATE_x (gas.f:172-173) and it stays on
 * the host in v1 of the boundary ("next" row N1 moves it to
 * the device).  A Fortran caller that keeps its own Update_RCONST_x does not need
 * this library; the synthetic-ensemble generator and the tests do.
 *
 * Inputs mirror the COMMON blocks Update_RCONST_x reads, one row per cell.
 */
#ifndef MISTRA_RCONST_H
#define MISTRA_RCONST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MISTRA_NPHRXN 47 /* global_params.f90: nphrxn */

typedef struct mistra_rate_inputs {
  int64_t ncell;
  const double *cb1;    /* [ncell][4]  COMMON /cb_1/ aircc, te, h2oppm, pk (kpp.f90:4315-4321) */
  const double *scal;   /* [ncell][13] conv1, xhal, xiod, xhet1, xhet2, xliq1..4, cvv1..4       */
  const double *ph_rat; /* [ncell][47] COMMON /ph_r_x/ (kpp.f90:4350-4362)                      */
  const double *conc;   /* [ncell][NSPEC] C = VAR|FIX at the start of the step                  */
  /* COMMON /kpp_rate_x/ arrays, KPP-species indexed; NULL = all zero */
  const double *yhenry; /* [ncell][NSPEC]      */
  const double *yxkmt;  /* [ncell][nkc][NSPEC] nkc = 2 (aer) or 4 (tot) */
  const double *ykef;   /* [ncell][nkc][NSPEC] */
  const double *ykeb;   /* [ncell][nkc][NSPEC] */
  const double *yxkmtd; /* [ncell][2][NSPEC]   */
  const double *yxeq;   /* [ncell][NSPEC]      */
  const double *ycw;    /* [ncell][nkc]        */
  const double *ycwd;   /* [ncell][2]          */
  int32_t f32_literals; /* 1: default-REAL literals binary32 (reference's preferred flags) */
  int32_t reserved;
} mistra_rate_inputs;

/* rconst [ncell][NREACT] out.  nthreads <= 1: serial. Returns 0 or -1 (bad args). */
int mistra_rconst_update(int mech, const mistra_rate_inputs *in, double *rconst, int nthreads);

/* 0-based KPP index of a species name in the mechanism (SPC_NAMES), -1 if absent. */
int mistra_rconst_spc_index(int mech, const char *name);

#ifdef __cplusplus
}
#endif
#endif
