/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for who may load this).
 *
 * Restatement of SUBROUTINE difc (/root/reference/src/str.f90:3271-3445): implicit turbulent exchange
 * of the chemical species of one column plus explicit subsidence, statement by statement; columns are
 * independent.  Parity unpinned by the reference (no tests or fixtures there).
 * Arrays as in include/mistra_difc.h; Fortran level k = index k-1. */
#include <stdint.h>
#include <stdlib.h>

void difc_oracle(int64_t ncol, int n, double dt, const double *atkh, const double *w, const double *am3,
                 const double *detw, const double *deta, int nfield, double **fs, const int32_t *row,
                 const int32_t *nproc)
{
  const int nm = n - 1;
#pragma omp parallel
  {
    double *xa = malloc(sizeof(double) * n * 7), *xb = xa + n, *xc = xb + n, *xd = xc + n, *xe = xd + n,
           *xf = xe + n, *c = xf + n;
#pragma omp for schedule(static)
    for (int64_t col = 0; col < ncol; ++col) {
      const double *ak = atkh + col * n, *wk = w + col * n, *am = am3 + col * n;
      /* indices below are Fortran's k (1-based) shifted by one */
      xa[0] = ak[0] * dt / (detw[0] * deta[0]);                       /* str.f90:3334 */
      xe[0] = 0.0;
      for (int k = 1; k < nm; ++k) {                                  /* do k=2,nm : 3336-3343 */
        xa[k] = ak[k] * dt / (detw[k] * deta[k]);
        xc[k] = xa[k - 1] * detw[k - 1] / detw[k];
        xb[k] = 1.0 + xa[k] + xc[k];
        xd[k] = xb[k] - xc[k] * xe[k - 1];
        xe[k] = xa[k] / xd[k];
        c[k] = wk[k] * dt / deta[k];
      }
      for (int f = 0; f < nfield; ++f) {
        double *s = fs[f] + col * (int64_t)n * row[f];
        const int r = row[f];
        for (int j = 0; j < nproc[f]; ++j) {                          /* 3346-3357 and the three copies below it */
          xf[0] = s[1 * r + j] / am[1];
          for (int k = 1; k < nm; ++k) xf[k] = (s[k * r + j] / am[k] + xc[k] * xf[k - 1]) / xd[k];
          for (int k = nm - 1; k >= 1; --k) s[k * r + j] = (xe[k] * s[(k + 1) * r + j] / am[k + 1] + xf[k]) * am[k];
          for (int k = 1; k < nm; ++k) s[k * r + j] = s[k * r + j] - c[k] * (s[(k + 1) * r + j] - s[k * r + j]);
        }
      }
    }
    free(xa);
  }
}

/* SUBROUTINE difp (/root/reference/src/str.f90:3137-3265): the same exchange for the particle spectrum,
 * on ff / rho, then subsidence, then fsum.  ff [ncol][n][row], rho [ncol][n], fsum [ncol][n]. */
void difp_oracle(int64_t ncol, int n, int row, double dt, const double *atkh, const double *w, const double *rho,
                 const double *detw, const double *deta, double *ff, double *fsum)
{
  const int nm = n - 1;
#pragma omp parallel
  {
    double *xa = malloc(sizeof(double) * n * 6), *xb = xa + n, *xc = xb + n, *xd = xc + n, *xe = xd + n, *c = xe + n;
    double *xf = malloc(sizeof(double) * (size_t)n * row);
#pragma omp for schedule(static)
    for (int64_t col = 0; col < ncol; ++col) {
      const double *ak = atkh + col * n, *wk = w + col * n, *rh = rho + col * n;
      double *f = ff + col * (int64_t)n * row;
      for (int k = 1; k <= nm; ++k)                                   /* do k=2,nm+1 : 3210-3212 */
        for (int j = 0; j < row; ++j) f[k * row + j] = f[k * row + j] / rh[k];
      xa[0] = ak[0] * dt / (detw[0] * deta[0]);                       /* 3215-3225 */
      xe[0] = 0.0;
      for (int k = 1; k < nm; ++k) {
        xa[k] = ak[k] * dt / (detw[k] * deta[k]);
        xc[k] = xa[k - 1] * detw[k - 1] / detw[k];
        xb[k] = 1.0 + xa[k] + xc[k];
        xd[k] = xb[k] - xc[k] * xe[k - 1];
        xe[k] = xa[k] / xd[k];
        c[k] = wk[k] * dt / deta[k];
      }
      for (int j = 0; j < row; ++j) xf[j] = f[1 * row + j];           /* xf(:,:,1) = ff(:,:,2) */
      for (int k = 1; k < nm; ++k)                                    /* 3230-3232 */
        for (int j = 0; j < row; ++j) xf[k * row + j] = (f[k * row + j] + xc[k] * xf[(k - 1) * row + j]) / xd[k];
      for (int k = nm - 1; k >= 1; --k)                               /* 3234-3236 */
        for (int j = 0; j < row; ++j) f[k * row + j] = xe[k] * f[(k + 1) * row + j] + xf[k * row + j];
      for (int k = 1; k <= nm; ++k)                                   /* 3238-3240 */
        for (int j = 0; j < row; ++j) f[k * row + j] = f[k * row + j] * rh[k];
      for (int k = 1; k < nm; ++k)                                    /* 3243-3246 */
        for (int j = 0; j < row; ++j) f[k * row + j] = f[k * row + j] - c[k] * (f[(k + 1) * row + j] - f[k * row + j]);
      for (int k = 1; k < n; ++k) {                                   /* 3248-3255 */
        double s = 0.0;
        for (int j = 0; j < row; ++j) s = s + f[k * row + j];
        fsum[col * n + k] = s;
      }
    }
    free(xa);
    free(xf);
  }
}
