/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for the rules).
 *
 * Plain-C restatement of the condensation / evaporation step on the 2-D particle grid
 * for one humid model layer: SUBROUTINE subkon (/root/reference/src/str.f90:4987-5204)
 * and the positive-definite advection scheme it calls, SUBROUTINE advec
 * (str.f90:5321-5516, Bott's area-preserving flux form), with the helper functions
 * diff_wat_vap (5210-5269), therm_conduct_air (5275-5315), xl21 (7640), p21 (7672).
 * Statement order, loop order and every floating-point association follow the Fortran.
 * Called in the reference from SUBROUTINE kon for every layer k = 2..nf+1 with
 * feu(k) >= 0.7 (str.f90:4615-4721).
 *
 * PARITY STATUS: "parity unpinned" - no tests/fixtures in the reference, no Fortran
 * compiler here; pinned by properties (tests/test_kon_oracle.py): advec conserves the
 * row sum and positivity, is the identity for u = 0, moves mass by the Courant number;
 * subkon closes the water budget (vapour lost = liquid gained) and converges.
 *
 * Layout: ffk [ncell][nka][nkt] = ffk(nkt,nka); rw [nka][nkt] = rw(nkt,nka);
 * qabs [3][nka][nkt][18] = qabs(18,nkt,nka,jptaerrad); totr [ncell][18].
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MB 18

typedef struct {
  int nka, nkt;
  double a0m, dlne;
  const double *en, *rn, *b0m;  /* [nka] */
  const double *ew, *e, *dew;   /* [nkt] */
  const double *rw;             /* [nka][nkt] */
  const double *qabs;           /* [3][nka][nkt][18] */
  const int *kw;                /* [nka] (layers only) */
  const double *rq;             /* [nka][nkt] (layers only) */
  int ka, reserved;
} kon_grid;

/* constants.f90:48-83 */
static const double gas_const = 8.3144743, M_air = 28.96546e-3, M_wat = 18.01528e-3;
static const double cp = 1005.0, rhow = 1000.0, pi = 3.1415926535897932;

static double xl21(double t) { return 3138708. + (-2339.4) * t; }                       /* str.f90:7661 */
static double p21(double t) { return 610.7 * exp(17.15 * (t - 273.15) / (t - 38.33)); } /* str.f90:7691 */
static double therm_conduct_air(double t) { return 4.39e-3 + 7.1e-5 * t; }              /* str.f90:5313 */
/* str.f90:5255-5257; cst2 = cst*P0/(T0**exponent) is a compile-time constant there:
 * 0.211e-4*101325/273.15**1.94, folded here to the same double */
static double diff_wat_vap(double t, double p)
{
  const double cst2 = 0.211e-4 * 101325. / 53286.64010226011;  /* 273.15**1.94 */
  return cst2 * pow(t, 1.94) / p;
}

/* SUBROUTINE advec (str.f90:5321-5516).  Returns 0, or 1 where the reference aborts
 * ("error with k_high or k_low"). */
int kon_oracle_advec(int nkt, double dt, const double *u, double *y)
{
  const double ymin = 1.e-32;
  double z[256];
  memcpy(z, y, sizeof(double) * nkt);
  for (int i = 0; i < nkt; ++i) y[i] = 0.0;
  int i0 = 1;
  while (z[i0 - 1] < ymin) {
    if (i0 == nkt) return 0;
    i0 = i0 + 1;
  }
  int i1 = nkt;
  while (z[i1 - 1] < ymin) i1 = i1 - 1;
  for (int i = i0; i <= i1; ++i) {
    if (z[i - 1] < ymin) continue;
    int k2 = 0, k1, k = i;
    double dt0, dt1 = dt, x0;
    if (fabs(u[k - 1]) > 0.0) dt0 = fmin(1.0 / fabs(u[k - 1]), dt1);
    else { y[k - 1] = y[k - 1] + z[i - 1]; continue; }
    x0 = (double)k + u[k - 1] * dt0;
    dt1 = dt1 - dt0;
    k1 = k;
    int done = 0;
    while (dt1 > 1.e-7) {
      if (u[k - 1] < 0.0) k = k - 1; else k = k + 1;
      if (k == k2) { y[k - 1] = y[k - 1] + z[i - 1]; done = 1; break; }
      k2 = k1;
      k1 = k;
      if (k < 1 || k > nkt) return 1;  /* the Fortran would index out of bounds here */
      if (fabs(u[k - 1]) > 0.0) dt0 = fmin(1.0 / fabs(u[k - 1]), dt1);
      else { y[k - 1] = y[k - 1] + z[i - 1]; done = 1; break; }
      x0 = (double)k + u[k - 1] * dt0;
      dt1 = dt1 - dt0;
    }
    if (done) continue;
    const int k_low = (int)floor(x0), k_high = k_low + 1;
    const double c0 = x0 - (double)k_low;
    if (k_low < 1 || (k_high > nkt && c0 > 0.0)) return 1;
    if (c0 > 0.0) {
      double x1;
      if (i == 1 || i == nkt) {
        x1 = c0 * z[i - 1];
      } else if (i == 2 || i == nkt - 1) {
        const double al = 1.0 - 2.0 * c0, al2 = al * al;
        const double a0 = (26.0 * z[i - 1] - z[i] - z[i - 2]) / 24.0;
        const double a1 = (z[i] - z[i - 2]) / 16.0;
        const double a2 = (z[i] + z[i - 2] - 2.0 * z[i - 1]) / 48.0;
        x1 = fmin(z[i - 1], a0 * c0 + a1 * (1.0 - al2) + a2 * (1.0 - al2 * al));
      } else {
        const double al = 1.0 - 2.0 * c0, al2 = al * al, al3 = al2 * al;
        const double zp2 = z[i + 1], zp1 = z[i], z0 = z[i - 1], zm1 = z[i - 2], zm2 = z[i - 3];
        const double a0 = (9.0 * (zp2 + zm2) - 116.0 * (zp1 + zm1) + 2134.0 * z0) / 1920.0;
        const double a1 = (-5.0 * (zp2 - zm2) + 34.0 * (zp1 - zm1)) / 384.0;
        const double a2 = (-zp2 + 12.0 * (zp1 + zm1) - 22.0 * z0 - zm2) / 384.0;
        const double a3 = (zp2 - 2.0 * (zp1 - zm1) - zm2) / 768.0;
        const double a4 = (zp2 - 4.0 * (zp1 + zm1) + 6.0 * z0 + zm2) / 3840.0;
        x1 = fmin(z0, a0 * c0 + a1 * (1.0 - al2) + a2 * (1.0 - al3) + a3 * (1.0 - al2 * al2)
                          + a4 * (1.0 - al2 * al3));
      }
      x1 = fmax(0.0, x1);
      y[k_low - 1] = y[k_low - 1] + z[i - 1] - x1;
      y[k_high - 1] = y[k_high - 1] + x1;
    } else {
      y[k_low - 1] = y[k_low - 1] + z[i - 1];
    }
  }
  return 0;
}

/* SUBROUTINE subkon for every layer.  status[c] = number of iterations used (1..10),
 * -1 = "no convergence of condensation iteration" (the reference warns and carries on),
 * -2 = advec aborted. */
void kon_oracle_subkon(const kon_grid *g, int64_t ncell, double dt, double *ffk_all,
                       const double *totr_all, const double *dfdt, const double *feualt_a,
                       const double *pp_a, double *to_a, const double *tn_a, double *xm1o_a,
                       const double *xm1n_a, const int32_t *kr_a, int32_t *status)
{
  const int nka = g->nka, nkt = g->nkt;
  const double r0 = gas_const / M_air, r1 = gas_const / M_wat;
  const size_t tile = (size_t)nka * nkt;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t c = 0; c < ncell; ++c) {
    double *ffk = ffk_all + c * tile;
    const double *totr = totr_all + c * MB;
    double to = to_a[c], xm1o = xm1o_a[c];
    const double tn = tn_a[c], xm1n = xm1n_a[c], pp = pp_a[c], feualt = feualt_a[c];
    double cd[128 * 128], cr[128 * 128], sr[128 * 128], falt[128 * 128];
    double cc[128], psi[128], u[128];
    const double zxl21 = xl21(to);
    const double xldcp = zxl21 / cp;
    const double xka = therm_conduct_air(to);
    const double xdv = diff_wat_vap(to, pp);
    const double xl = 24.483 * to / pp;
    const double deltav = 1.3 * xl, deltat = 2.7 * xl;
    const double rho = pp / (r0 * to * (1.0 + 0.61 * xm1o));
    const double rho21 = p21(to) / (r1 * to);
    const double rho21s = (zxl21 / (r1 * to) - 1.0) * rho21 / to;
    const double a0 = g->a0m / to;
    const double xdv0 = xdv * sqrt(2.0 * pi / (r1 * to)) / 3.6e-08;
    const double xka0 = xka * sqrt(2.0 * pi / (r0 * to)) / (7.e-07 * rho * cp);
    int kr0 = kr_a[c];
    const int ib0 = (totr[0] < 1.0) ? 7 : 1;
    for (int ia = 1; ia <= nka; ++ia)
      for (int jt = 1; jt <= nkt; ++jt) {
        const int jtp = jt + 1 < nkt ? jt + 1 : nkt;
        const double de0 = g->dew[jt - 1], dep = g->dew[jtp - 1], de0p = de0 + dep;
        const double rk = g->rw[(ia - 1) * nkt + jt - 1];
        const int q = (ia - 1) * nkt + jt - 1;
        sr[q] = fmax(0.1, exp(a0 / rk - g->b0m[ia - 1] * g->en[ia - 1] / g->ew[jt - 1]));
        const double xdvs = xdv / (rk / (rk + deltav) + xdv0 / rk);
        const double xkas = xka / (rk / (rk + deltat) + xka0 / rk);
        const double x1 = rhow * (zxl21 + xkas / (xdvs * rho21s * sr[q]));
        cd[q] = 3.e12 * rho21 * xkas / (x1 * rk * rk * rho21s * sr[q]);
        if (kr0 == 3 && g->rn[ia - 1] < 0.5) kr0 = 2;     /* sticks, as in the reference (str.f90:5141) */
        double rad = 0.0;
        for (int ib = ib0; ib <= MB; ++ib) {
          const double *qa = g->qabs + ((size_t)(kr0 - 1) * nka + (ia - 1)) * nkt * MB;
          rad = rad + totr[ib - 1] * (qa[(jt - 1) * MB + ib - 1] * de0 + qa[(jtp - 1) * MB + ib - 1] * dep) / de0p;
        }
        cr[q] = rad * 7.5e5 / (rk * x1) - rhow * 4190. * (tn - to) / (dt * x1);
      }
    memcpy(falt, ffk, sizeof(double) * tile);
    double feuneu = feualt + dfdt[c] * dt;
    if (feualt < 0.95) feuneu = xm1n * pp / (p21(tn) * (.62198 + .37802 * xm1n));
    double fquer = 0.5 * (feuneu + feualt);
    double res = 0.0, fqa = 0.0;
    const double aa0 = 1.0 / dt;
    int st = -1;
    for (int itk = 1; itk <= 10; ++itk) {
      double dwsum = 0.0;
      int bad = 0;
      for (int ia = 1; ia <= nka && !bad; ++ia) {
        const int r = (ia - 1) * nkt;
        for (int jt = 1; jt <= nkt; ++jt) {
          psi[jt - 1] = falt[r + jt - 1];
          cc[jt - 1] = (cd[r + jt - 1] * (fquer - sr[r + jt - 1]) - cr[r + jt - 1]) / g->dlne;
        }
        u[0] = fmax(0.0, cc[0]);
        for (int jt = 2; jt <= nkt - 1; ++jt)
          u[jt - 1] = 0.5 * (cc[jt - 1] + fabs(cc[jt - 1]) + cc[jt - 2] - fabs(cc[jt - 2]));
        u[nkt - 1] = fmin(0.0, cc[nkt - 2]);
        if (kon_oracle_advec(nkt, dt, u, psi)) { bad = 1; break; }
        for (int jt = 1; jt <= nkt; ++jt) {
          ffk[r + jt - 1] = psi[jt - 1];
          dwsum = dwsum + (psi[jt - 1] - falt[r + jt - 1]) * g->e[jt - 1];
        }
      }
      if (bad) { st = -2; break; }
      const double dmsum = dwsum / rho;
      const double dtsum = xldcp * dmsum;
      xm1o = xm1n - dmsum;
      to = tn + dtsum;
      const double p1 = xm1o * pp / (0.62198 + 0.37802 * xm1o);
      feuneu = p1 / p21(to);
      const double resold = res;
      res = feuneu + feualt - 2.0 * fquer;
      if (fabs(res) < 1.e-6) { st = itk; break; }
      const double dres = res - resold;
      double aa = aa0;
      if (itk > 1 && fabs(dres) > 1.e-8) aa = (fqa - fquer) / dres;
      fqa = fquer;
      fquer = fquer + aa * res;
    }
    to_a[c] = to;
    xm1o_a[c] = xm1o;
    status[c] = st;
  }
}

/* FUNCTION rgl (str.f90:2164-2251): Koehler equilibrium radius by Newton iteration */
static double rgl(double r_dry, double a, double b, double feu)
{
  if (feu >= 1.0) return r_dry;
  const double zlogf = log(feu);
  const double alpha = a / r_dry;
  double xalt = exp(feu), xneu = xalt;
  for (int ij = 1; ij <= 100; ++ij) {
    const double falt = (xalt * xalt * xalt - 1.0) * (xalt * zlogf - alpha) + b * xalt;
    const double fstralt = (4.0 * (xalt * xalt * xalt) - 1.0) * zlogf - 3.0 * (xalt * xalt) * alpha + b;
    xneu = xalt - falt / fstralt;
    if (fabs(xneu - xalt) < 1.e-7 * xalt) break;
    xalt = xneu;
  }
  return r_dry * xneu;
}

typedef struct {
  double *ff, *t, *talt, *xm1, *xm1a, *feu, *dfddt, *xm2, *dtcon;
  const double *p, *totrad;
  const int32_t *nar;
  double *vol1_a, *vol1_d, *part_o_a, *part_o_d, *part_n_a, *part_n_d, *vol2, *pntot;
  int32_t *status;
} kon_state;

/* The layer loop of SUBROUTINE kon (str.f90:4615-4772) for ncell layers: bin sums before,
 * dry branch (feu < 0.7: SUBROUTINE equil case 1, str.f90:4801-4981) or humid branch
 * (subkon + write-back 4708-4721), bin sums after.  status: 0 dry, else as subkon. */
void kon_oracle_layers(const kon_grid *g, int64_t ncell, double dt, int chem, const kon_state *s)
{
  const int nka = g->nka, nkt = g->nkt, ka = g->ka;
  const size_t tile = (size_t)nka * nkt;
  const double z4pi3 = 4.0 * pi / 3.0;               /* str.f90:4560 */
  for (int64_t c = 0; c < ncell; ++c) {
    double *ffk = s->ff + c * tile;
    const double tn = s->t[c], xm1n = s->xm1[c], pp = s->p[c];
    s->dtcon[c] = 0.0;
    if (chem) {                                      /* str.f90:4625-4659 */
      double v2[4] = {0, 0, 0, 0};
      for (int ia = 0; ia < nka; ++ia) {
        double va = 0, pa = 0, vd = 0, pd = 0;
        for (int jt = 0; jt < nkt; ++jt) {
          const double f = ffk[ia * nkt + jt], r = g->rq[ia * nkt + jt];
          const double v = f * z4pi3 * (r * r * r);
          if (jt < g->kw[ia]) { va = va + v; pa = pa + f; } else { vd = vd + v; pd = pd + f; }
        }
        s->vol1_a[c * nka + ia] = va; s->part_o_a[c * nka + ia] = pa;
        s->vol1_d[c * nka + ia] = vd; s->part_o_d[c * nka + ia] = pd;
        if (ia < ka) { v2[0] = v2[0] + va; v2[2] = v2[2] + vd; } else { v2[1] = v2[1] + va; v2[3] = v2[3] + vd; }
      }
      for (int k = 0; k < 4; ++k) s->vol2[c * 4 + k] = v2[k];
    }
    if (s->feu[c] < 0.7) {                           /* str.f90:4663-4672 */
      const double feun = xm1n * pp / ((0.62198 + 0.37802 * xm1n) * p21(tn));
      s->feu[c] = feun;
      const double a0 = g->a0m / tn;
      double x2 = 0.0;
      for (int ia = 0; ia < nka; ++ia) {             /* equil case 1 */
        double tot = 0.0;
        for (int jt = 0; jt < nkt; ++jt) { tot = tot + ffk[ia * nkt + jt]; ffk[ia * nkt + jt] = 0.0; }
        const double rn = g->rn[ia];
        const double rg = rgl(rn, a0, g->b0m[ia] * 2.0, feun);
        const double eg = (4.e-09 * pi / 3.0) * (rg * rg * rg - rn * rn * rn);
        int jt = 1;
        while (jt < nkt && eg > g->ew[jt - 1]) jt = jt + 1;
        ffk[ia * nkt + jt - 1] = tot;
      }
      for (int ia = 0; ia < nka; ++ia)
        for (int jt = 0; jt < nkt; ++jt) x2 = x2 + ffk[ia * nkt + jt] * g->e[jt];
      s->xm2[c] = x2;
      if (s->status) s->status[c] = 0;
    } else {                                         /* str.f90:4676-4721 */
      double to = s->talt[c], xm1o = s->xm1a[c];
      const double feualt = s->feu[c];
      int32_t st, kr = s->nar[c];
      kon_oracle_subkon(g, 1, dt, ffk, s->totrad + c * MB, &s->dfddt[c], &feualt, &pp, &to, &tn, &xm1o, &xm1n,
                        &kr, &st);
      s->t[c] = to; s->talt[c] = to; s->xm1[c] = xm1o; s->xm1a[c] = xm1o;
      s->feu[c] = xm1o * pp / ((0.62198 + 0.37802 * xm1o) * p21(to));
      s->dfddt[c] = (s->feu[c] - feualt) / dt;
      double x2 = 0.0;
      for (int ia = 0; ia < nka; ++ia)
        for (int jt = 0; jt < nkt; ++jt) x2 = x2 + ffk[ia * nkt + jt] * g->e[jt];
      s->xm2[c] = x2;
      s->dtcon[c] = (to - tn) / dt;
      if (s->status) s->status[c] = st;
    }
    if (chem) {                                      /* str.f90:4724-4770 */
      double pn[4] = {0, 0, 0, 0};
      for (int ia = 0; ia < nka; ++ia) {
        double pa = 0, pd = 0;
        for (int jt = 0; jt < nkt; ++jt) {
          const double f = ffk[ia * nkt + jt];
          if (jt < g->kw[ia]) pa = pa + f; else pd = pd + f;
        }
        s->part_n_a[c * nka + ia] = pa; s->part_n_d[c * nka + ia] = pd;
        if (ia < ka) { pn[0] = pn[0] + pa; pn[2] = pn[2] + pd; } else { pn[1] = pn[1] + pa; pn[3] = pn[3] + pd; }
      }
      for (int k = 0; k < 4; ++k) s->pntot[c * 4 + k] = pn[k];
    }
  }
}
