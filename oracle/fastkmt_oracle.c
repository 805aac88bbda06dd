/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for who may load this).
 *
 * Restatement of SUBROUTINE fast_k_mt_a (/root/reference/src/kpp.f90:2683-2947) = fast_k_mt_t
 * (kpp.f90:2421-2676, same body) and of FUNCTION vterm (str.f90:2793-2864): mass-transfer
 * coefficients of the exchanged species and sedimentation velocity of each chemistry bin, integrated
 * over the 2-D particle spectrum.  Statement and summation order of the reference (one running sum
 * per (bin, species) over ia, then jt).  Parity unpinned by the reference (no tests or fixtures
 * there).  Arrays as in include/mistra_fastkmt.h. */
#include <math.h>
#include <stdint.h>

static double vterm_oracle(double a, double t, double p)       /* str.f90:2793-2864 */
{
  const double g = 9.80665, gas_const = 8.3144743, M_air = 28.96546e-3;   /* constants.f90:48-79 */
  const double r0 = gas_const / M_air, rhow = 1000.0;
  const double b0 = -.318657e+1, b1 = .992696e+0, b2 = -.153193e-2, b3 = -.987059e-3, b4 = -.578878e-3,
               b5 = +.855176e-4, b6 = -.327815e-5;
  const double c1 = 2.0 * g / 9.0, c2 = 1.26, P0 = 101325, T0 = 293.15, lambda0 = 6.6e-8;
  const double c3 = c2 * lambda0 * P0 / T0, c4 = 32.0 * g / 3.0;
  const double rho_a = p / (r0 * t);
  const double eta = 3.7957e-06 + 4.9e-08 * t;
  if (a <= 1.e-5) return c1 * a * a * (rhow - rho_a) / eta * (1.0 + c3 * t / (a * p));
  const double best = c4 * (a * a * a) * (rhow - rho_a) * rho_a / (eta * eta);
  const double x = log(best);
  double y = b6 * x + b5;
  y = y * x + b4;
  y = y * x + b3;
  y = y * x + b2;
  y = y * x + b1;
  y = y * x + b0;
  return eta * exp(y) / (2. * rho_a * a);
}

void vterm_oracle_vec(int64_t n, const double *a, const double *t, const double *p, double *out)
{
  for (int64_t i = 0; i < n; ++i) out[i] = vterm_oracle(a[i], t[i], p[i]);
}

void fastkmt_oracle(int64_t ncell, int nka, int nkt, int ka, int ial, int nkc, int nkc_l, int nspec, int nx,
                    const int32_t *lex, const int32_t *kw, const double *rq, const double *ff,
                    const double *freep, const double *t, const double *p, const double *cw, const double *cm,
                    const double *alpha, const double *vmean, double *xkmt, double *vt)
{
  const double pi = 3.1415926535897932;             /* constants.f90:54 */
  const double z4pi3 = 4.0 * pi / 3.0;              /* kpp.f90:2758 */
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t k = 0; k < ncell; ++k) {
    const double *f = ff + k * nka * nkt;
    for (int kc = 0; kc < nkc_l; ++kc) {
      const int llchem = cm[k * nkc + kc] > 0.0;    /* kpp.f90:2826-2833 */
      const int lmax = llchem ? nx : 1;
      double x1 = 0.0, xk1 = 0.0, xx1 = 0.0;
      for (int l = 0; l < lmax; ++l) {
        int ia0, ia1;                               /* summation limits (1), 2869-2881; 0-based, exclusive end */
        if (kc == 0 || kc == 2) { ia0 = ial - 1; ia1 = ka; }
        else                    { ia0 = ka;      ia1 = nka; }
        if (llchem) { x1 = 0.0; xk1 = 0.0; }
        if (l == 0) xx1 = 0.0;
        const double al = alpha[k * nspec + lex[l] - 1], vm = vmean[k * nspec + lex[l] - 1];
        if (al > 0.0) x1 = 4. / (3. * al);
        for (int ia = ia0; ia < ia1; ++ia) {
          int jt0, jt1;                             /* summation limits (2), 2894-2901 */
          if (kc < 2) { jt0 = 0;      jt1 = kw[ia]; }
          else        { jt0 = kw[ia]; jt1 = nkt; }
          for (int jt = jt0; jt < jt1; ++jt) {
            const double rqq = rq[ia * nkt + jt] * 1.e-6;            /* rqm, 2806 */
            if (llchem) {
              const double x2 = vm / (rqq / freep[k] + x1);          /* 2915 */
              xk1 = xk1 + x2 * rqq * rqq * f[ia * nkt + jt] * 1.e6;  /* 2921 */
            }
            if (l == 0) {
              const double xvs = vterm_oracle(rqq, t[k], p[k]);
              xx1 = xx1 + rqq * rqq * rqq * xvs * f[ia * nkt + jt] * 1.e6;   /* 2926 */
            }
          }
        }
        if (cw[k * nkc + kc] > 0.0) {               /* 2932-2937 */
          if (llchem) xkmt[(k * nkc + kc) * nspec + lex[l] - 1] = z4pi3 / cw[k * nkc + kc] * xk1;
          if (l == 0) vt[k * nkc + kc] = z4pi3 / cw[k * nkc + kc] * xx1;
        }
      }
    }
  }
}
