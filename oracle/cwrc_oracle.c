/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for who may load this).
 *
 * Restatement of SUBROUTINE cw_rc (/root/reference/src/kpp.f90:2152-2414): liquid water content,
 * mean radius and water volume of the four chemistry bins from the 2-D particle spectrum, and the
 * chemistry switches cm / conv2.  Statement and summation order of the reference (one running sum
 * per bin over ia, then jt).  Parity unpinned by the reference (no tests or fixtures there).
 * Arrays as in include/mistra_cwrc.h. */
#include <math.h>
#include <stdint.h>

void cwrc_oracle(int64_t ncell, int nka, int nkt, int ka, int ial, double xcryssulf, double xcrysss,
                 double xdelisulf, double xdeliss, const int32_t *kw, const double *e, const double *rq,
                 const double *ff, const double *feu, const int32_t *cloud, double *rc, double *cw, double *cm,
                 double *conv2)
{
  const double pi = 3.1415926535897932;             /* constants.f90 */
  const double xpi = 4.0 / 3.0 * pi;                /* kpp.f90:2204 */
  const double cwm = 1.e-1, cwmd = 1.e2;            /* kpp.f90:2207-2208 */
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < ncell; ++k) {
    const double *f = ff + k * nka * nkt;
    double s_cw[4] = {0, 0, 0, 0}, s_rc[4] = {0, 0, 0, 0}, s_cm[4] = {0, 0, 0, 0};
    for (int ia = ial - 1; ia < nka; ++ia) {        /* kpp.f90:2285-2322: small classes, then large ones */
      const int ba = ia < ka ? 0 : 1;
      for (int jt = 0; jt < nkt; ++jt) {
        const int b = jt < kw[ia] ? ba : ba + 2;
        const double r = rq[ia * nkt + jt];
        const double x0 = f[ia * nkt + jt] * xpi * (r * r * r);
        s_cw[b] = s_cw[b] + x0;
        s_rc[b] = s_rc[b] + x0 * r;
        const double x1 = f[ia * nkt + jt] * e[jt];
        s_cm[b] = s_cm[b] + x1;
      }
    }
    for (int b = 0; b < 4; ++b) {
      rc[k * 4 + b] = s_cw[b] > 0.0 ? s_rc[b] / s_cw[b] * 1.e-6 : 0.0;   /* 2335-2354 */
      cw[k * 4 + b] = s_cw[b] * 1.e-12;                                  /* 2356-2359 */
    }
    int on[4] = {0, 0, 0, 0};
    if (!(feu[k] < fmin(xcryssulf, xcrysss))) {                          /* 2367-2410 */
      on[0] = s_cw[0] >= cwm && ((cloud[k * 4 + 0] && feu[k] >= xcryssulf) || feu[k] >= xdelisulf);
      on[1] = s_cw[1] >= cwm && ((cloud[k * 4 + 1] && feu[k] >= xcrysss) || feu[k] >= xdeliss);
      on[2] = s_cw[2] >= cwmd;
      on[3] = s_cw[3] >= cwmd;
    }
    for (int b = 0; b < 4; ++b) {
      cm[k * 4 + b] = on[b] ? s_cm[b] * 1.e-3 : 0.0;
      conv2[k * 4 + b] = on[b] ? 1.e9 / s_cw[b] : 0.0;
    }
  }
}
