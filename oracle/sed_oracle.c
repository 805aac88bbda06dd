/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for who may load this).
 *
 * Restatement of the gravitational-settling operators of the reference, statement by statement:
 *   SUBROUTINE sedp    /root/reference/src/str.f90:2257-2411  particle spectrum ff
 *   SUBROUTINE sedc    /root/reference/src/str.f90:2417-2621  dry deposition / emission of gases (the loop 2567-2596)
 *   SUBROUTINE sedl    /root/reference/src/str.f90:2627-2787  aqueous species sl1, sion1
 *   SUBROUTINE advsed0 /root/reference/src/str.f90:5522-5579  upstream advection
 *   SUBROUTINE advsed1 /root/reference/src/str.f90:5585-5691  Bott's area-preserving flux form, 4th order
 *   FUNCTION   vterm   /root/reference/src/str.f90:2793-2864
 * Columns are independent.  The reference holds no tests or fixtures for these routines; this file is pinned by
 * tests/golden/make_sed_reference.py, which executes the reference's own Fortran statements (tests/test_sed_oracle.py).
 * Arrays as in include/mistra_sed.h; Fortran level k = index k-1. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double dmin(double a, double b) { return b < a ? b : a; }   /* gfortran's MIN / MAX: first argument on ties */
static double dmax(double a, double b) { return b > a ? b : a; }

double sed_vterm(double a, double t, double p)                    /* str.f90:2793-2864 */
{
  const double g = 9.80665, gas_const = 8.3144743, M_air = 28.96546e-3;   /* constants.f90 */
  const double r0 = gas_const / M_air, rhow = 1000.0;
  const double b0 = -.318657e+1, b1 = .992696e+0, b2 = -.153193e-2, b3 = -.987059e-3, b4 = -.578878e-3,
               b5 = +.855176e-4, b6 = -.327815e-5;
  const double c1 = 2.0 * g / 9.0, c2 = 1.26, P0 = 101325, T0 = 293.15, lambda0 = 6.6e-8;
  const double c3 = c2 * lambda0 * P0 / T0, c4 = 32.0 * g / 3.0;
  const double rho_a = p / (r0 * t);
  const double eta = 3.7957e-06 + 4.9e-08 * t;
  if (a <= 1.e-5) return c1 * a * a * (rhow - rho_a) / eta * (1.0 + c3 * t / (a * p));
  {
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
}

/* y, c: [nf], index 0 = level 1 */
void sed_advsed0(int nf, const double *c, double *y)              /* str.f90:5522-5579 */
{
  double *fm = malloc(sizeof(double) * 2 * nf), *fp = fm + nf;
  for (int i = 0; i < nf - 1; ++i) {
    fm[i] = -dmin(0.0, c[i]) * y[i + 1];
    fp[i] = dmax(0.0, c[i]) * y[i];
  }
  for (int i = 1; i < nf - 1; ++i) y[i] = y[i] - fm[i - 1] + fp[i - 1] + fm[i] - fp[i];
  free(fm);
}

void sed_advsed1(int nf, const double *c, double *y)              /* str.f90:5585-5691 */
{
  double *a0 = malloc(sizeof(double) * 6 * nf), *a1 = a0 + nf, *a2 = a1 + nf, *a3 = a2 + nf, *a4 = a3 + nf, *fm = a4 + nf;
  /* 0-based: Fortran index i -> i-1 */
  a0[1] = (26.0 * y[1] - y[2] - y[0]) / 24.0;
  a1[1] = (y[2] - y[0]) / 16.0;
  a2[1] = (y[2] + y[0] - 2.0 * y[1]) / 48.0;
  a3[1] = 0.0;
  a4[1] = 0.0;
  for (int i = 2; i <= nf - 3; ++i) {
    a0[i] = (9.0 * (y[i + 2] + y[i - 2]) - 116.0 * (y[i + 1] + y[i - 1]) + 2134.0 * y[i]) / 1920.0;
    a1[i] = (-5.0 * (y[i + 2] - y[i - 2]) + 34.0 * (y[i + 1] - y[i - 1])) / 384.0;
    a2[i] = (-y[i + 2] + 12.0 * (y[i + 1] + y[i - 1]) - 22.0 * y[i] - y[i - 2]) / 384.0;
    a3[i] = (y[i + 2] - 2.0 * (y[i + 1] - y[i - 1]) - y[i - 2]) / 768.0;
    a4[i] = (y[i + 2] - 4.0 * (y[i + 1] + y[i - 1]) + 6.0 * y[i] + y[i - 2]) / 3840.0;
  }
  {
    const int m = nf - 2;                                         /* Fortran nf-1 */
    a0[m] = (26.0 * y[m] - y[m + 1] - y[m - 1]) / 24.0;
    a1[m] = (y[m + 1] - y[m - 1]) / 16.0;
    a2[m] = (y[m + 1] + y[m - 1] - 2.0 * y[m]) / 48.0;
    a3[m] = 0.0;
    a4[m] = 0.0;
  }
  double cl = -c[nf - 2];
  fm[nf - 2] = dmin(y[nf - 1], cl * (y[nf - 1] - (1.0 - cl) * (y[nf - 1] - y[nf - 2]) * 0.5));
  double clm = cl;
  for (int i = nf - 2; i >= 1; --i) {
    cl = clm;
    clm = -c[i - 1];
    const double x1 = 1.0 - 2.0 * cl;
    const double x2 = x1 * x1;
    const double x3 = x1 * x2;
    const double ymin = dmin(y[i], y[i + 1]);
    const double ymax = dmax(y[i], y[i + 1]);
    double fmim = dmax(0.0, a0[i] * cl - a1[i] * (1.0 - x2) + a2[i] * (1.0 - x3) - a3[i] * (1.0 - x1 * x3)
                                + a4[i] * (1.0 - x2 * x3));
    fmim = dmin(fmim, y[i] - ymin + fm[i]);
    fmim = dmax(fmim, y[i] - ymax + fm[i]);
    fmim = dmax(0.0, fmim - (cl - clm) * y[i]);
    const double w = y[i] / dmax(fmim + 1.e-15, y[i]);
    fm[i - 1] = fmim * w;
  }
  y[0] = y[0] + fm[0];
  for (int i = 1; i < nf - 1; ++i) y[i] = y[i] - fm[i - 1] + fm[i];
  y[nf - 1] = y[nf - 1] - fm[nf - 2];
  free(a0);
}

/* SUBROUTINE sedp.  ff [ncol][n][nka][nkt]; t, p [ncol][n]; vd [ncol][nka][nkt]; rq [nka][nkt]; e [nkt]; kw [nka]
 * (1-based class indices as the reference holds them); diag [ncol][4] = ajs (out), trdep, ds1, ds2 (in/out).
 * x0 of the reference is a local that a class without particles leaves as the previous class left it
 * (str.f90:2352, 2397): reproduced; its value before the first class is taken as 0. */
void sedp_oracle(int64_t ncol, int n, int nf, int nka, int nkt, double dt, const double *detw, const double *deta,
                 const double *t, const double *p, const double *rq, const double *e, const int32_t *kw,
                 const double *vd, double *ff, double *diag)
{
  const int64_t row = (int64_t)nka * nkt;
#pragma omp parallel
  {
    double *c = malloc(sizeof(double) * 2 * nf), *psi = c + nf;
#pragma omp for schedule(static)
    for (int64_t col = 0; col < ncol; ++col) {
      const double *tk = t + col * n, *pk = p + col * n, *vdc = vd + col * row;
      double *f = ff + col * n * row, *dg = diag + col * 4;
      double ajs = 0.0, x0 = 0.0;
      c[nf - 1] = 0.0;
      const double x3 = -deta[1];
      for (int ia = 0; ia < nka; ++ia)
        for (int jt = 0; jt < nkt; ++jt) {
          const int64_t q = (int64_t)ia * nkt + jt;
          const double ww = -1. * sed_vterm(rq[q] * 1.e-6, tk[nf - 1], pk[nf - 1]);
          double dt0 = dt, xsum = 0.0;
          for (int k = 1; k < nf; ++k) {
            psi[k] = f[k * row + q] * detw[k];
            xsum = xsum + psi[k];
          }
          if (xsum > 1.e-6) {
            x0 = 0.0;
            while (dt0 > 0.1) {
              const double dtmax = dmin(dt0, x3 / (ww));
              for (int k = 1; k < nf; ++k) c[k] = dtmax / deta[k] * (-1. * sed_vterm(rq[q] * 1.e-6, tk[k], pk[k]));
              c[1] = dmin(c[1], dtmax / deta[1] * vdc[q] * (-1.));
              c[0] = c[1];
              dt0 = dt0 - dtmax;
              const double x1 = psi[1];
              psi[0] = x1;
              if (rq[q] < 1.0) sed_advsed0(nf, c, psi); else sed_advsed1(nf, c, psi);
              x0 = x0 + psi[0] - x1;
            }
            for (int k = 1; k < nf - 1; ++k) f[k * row + q] = psi[k] / detw[k];
            f[(nf - 1) * row + q] = f[(nf - 2) * row + q];
          }
          const double x2 = x0 * e[jt] * detw[1];
          ajs = ajs + x2 / dt;
          dg[1] = dg[1] + x2;
          if (jt + 1 <= kw[ia]) dg[2] = dg[2] + x2; else dg[3] = dg[3] + x2;
        }
      dg[0] = ajs;
    }
    free(c);
  }
}

/* SUBROUTINE sedl, both halves.  s [ncol][n][nkc][jx] (sl1 with jx = j2, sion1 with jx = j6): bins 0 .. nkc_l-1
 * settle; rc, vt [ncol][n][nkc] (vt(nkc,nf): levels above nf unused); vdm [ncol][nkc]. */
void sedl_oracle(int64_t ncol, int n, int nf, int nkc, int nkc_l, int jx, double dt, const double *detw,
                 const double *deta, const double *t, const double *p, const double *rc, const double *vt,
                 const double *vdm, double *s)
{
  const int64_t row = (int64_t)nkc * jx;
#pragma omp parallel
  {
    double *c = malloc(sizeof(double) * 3 * nf), *psi = c + nf, *cc = psi + nf;
#pragma omp for schedule(static)
    for (int64_t col = 0; col < ncol; ++col) {
      const double *tk = t + col * n, *pk = p + col * n, *rck = rc + col * n * nkc, *vtk = vt + col * n * nkc;
      double *sc = s + col * n * row;
      c[nf - 1] = 0.0;
      const double xfac = 1.e6;
      for (int kc = 0; kc < nkc_l; ++kc) {
        for (int k = 1; k < nf; ++k) {
          const double xxx = 0.01;
          const double x4 = dmax(xxx, xfac * rck[k * nkc + kc]);
          cc[k] = (-1.0 * sed_vterm(x4 * 1.e-6, tk[k], pk[k])) / deta[k];
          cc[k] = dmin(cc[k], -1.0 * vtk[k * nkc + kc] / deta[k]);
        }
        cc[1] = dmin(cc[1], -1.0 / deta[1] * vdm[col * nkc + kc]);
        for (int l = 0; l < jx; ++l) {
          double *sl = sc + (int64_t)kc * jx + l;
          for (int k = 1; k < nf; ++k) psi[k] = sl[k * row] * detw[k];
          double dt0 = dt, x0 = 0.0;
          const double xxxt = -.999 / cc[1];
          while (dt0 > 0.1) {
            const double dtmax = dmin(dt0, xxxt);
            dt0 = dt0 - dtmax;
            for (int k = 1; k < nf; ++k) c[k] = cc[k] * dtmax;
            c[0] = c[1];
            const double x1 = psi[1];
            psi[0] = x1;
            sed_advsed1(nf, c, psi);
            x0 = x0 + psi[0] - x1;
          }
          for (int k = 1; k < nf - 1; ++k) sl[k * row] = psi[k] / detw[k];
          sl[0] = sl[0] + x0 * deta[1];
        }
      }
    }
    free(c);
  }
}

/* SUBROUTINE sedc, the species loop str.f90:2567-2596 (x4 = 1): s1 [ncol][n][j1] levels 1 and 2; vg [j1] as the
 * statements above the loop leave it; es1 [j1]. */
void sedc_oracle(int64_t ncol, int n, int j1, double dt, const double *detw, const double *deta, const double *vg,
                 const double *es1, double *s1)
{
  const double Avogadro = 6.022140857e+23;                        /* constants.f90 */
  const double x4 = 1.0;
#pragma omp parallel for schedule(static)
  for (int64_t col = 0; col < ncol; ++col) {
    double *lev1 = s1 + col * n * j1, *lev2 = lev1 + j1;
    for (int j = 0; j < j1; ++j) {
      const double w = vg[j];
      if (w >= 1.e-5) {
        const double s12old = lev2[j];
        lev2[j] = lev2[j] * exp(-dt / deta[1] * vg[j]);
        lev1[j] = lev1[j] + (s12old - lev2[j]) * deta[1];
      }
      lev2[j] = lev2[j] + es1[j] * x4 * dt * 1.e+4 / (detw[1] * Avogadro);
    }
  }
}
