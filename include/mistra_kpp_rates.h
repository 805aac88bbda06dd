/* mistra_kpp_rates.h - one chemistry call with the rate constants formed ON THE DEVICE from compact inputs
 * (part of libmistra_kpp.so).
 *
 * In the reference, kpp_driver fills the COMMON blocks /cb_1/, /ph_r_x/, /kpp_rate_x/ for a layer
 * (/root/reference/src/kpp.f90:4315-4468), x_drive calls Update_RCONST_x and INTEGRATE_x (aer.f:216-217).  With
 * mistra_kpp_integrate the host still evaluates Update_RCONST_x and ships RCONST(NREACT) per cell - 60-80 % of the
 * bytes that cross PCIe.  Here the host ships what Update_RCONST_x READS instead, in compact form, and the library
 * runs  expand -> Update_RCONST_x (rconst_kernel, include/mistra_rconst_cuda.h) -> INTEGRATE_x  chunk by chunk on
 * the device:
 *   - the per-layer scalars cb1[4], scal[13], ph_rat[47], ycw[nkc], ycwd[2] as in include/mistra_rconst.h;
 *   - the KPP-species-indexed exchange arrays of /kpp_rate_x/ (yhenry, yxkmt, ykef, ykeb, yxkmtd, yxeq) only for
 *     the species that are ever non-zero: the ~50 exchanged species of lex (kpp.f90:2683) and the 4 dry-het
 *     species (kpp.f90:4697), given once as an index list per array;
 *   - the start-of-step concentrations Update_RCONST_x reads (aer.f:742,769,774) are VAR|FIX of the call itself.
 * Per aer cell that is ~0.6 k doubles instead of 1.24 k (RCONST alone is 979).
 */
#ifndef MISTRA_KPP_RATES_H
#define MISTRA_KPP_RATES_H
#include "mistra_kpp.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mistra_rate_list {
  int32_t n;          /* species carried (0: the whole array is zero, val / idx may be NULL) */
  int32_t reserved;
  const int32_t *idx; /* [n] 0-based KPP species indices (into C = VAR|FIX), ascending, the same for every cell */
  const double *val;  /* [ncell][nk][n]: nk = 1 (yhenry, yxeq), nkc (yxkmt, ykef, ykeb; 2 aer, 4 tot), 2 (yxkmtd) */
} mistra_rate_list;

typedef struct mistra_rate_inputs_compact {
  const double *cb1;    /* [ncell][4]  aircc, te, h2oppm, pk          (COMMON /cb_1/)        */
  const double *scal;   /* [ncell][13] conv1, xhal, xiod, xhet1, xhet2, xliq1..4, cvv1..4    */
  const double *ph_rat; /* [ncell][47]                                 (COMMON /ph_r_x/)      */
  const double *ycw;    /* [ncell][nkc] or NULL */
  const double *ycwd;   /* [ncell][2]   or NULL */
  mistra_rate_list yhenry, yxkmt, ykef, ykeb, yxkmtd, yxeq;
  int32_t f32_literals; /* as in mistra_rate_inputs */
  int32_t reserved;
} mistra_rate_inputs_compact;

/* Same contract as mistra_kpp_integrate (host buffers, var in/out, per-cell ierr / stats / hexit / texit, synchronous),
 * with `rates` in place of rconst.  Bytes moved host -> device per call are returned through *h2d_bytes if not NULL. */
int mistra_kpp_integrate_rates(int mech, int64_t ncell, const mistra_rate_inputs_compact *rates, const double *fix,
                               double *var, double t0, double t1, const mistra_kpp_opts *o, int32_t *ierr,
                               int32_t *stats, double *hexit, double *texit, int64_t *h2d_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif
