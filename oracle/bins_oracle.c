/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for the rules).
 *
 * Plain-C restatement of the 2-D aerosol/droplet bin redistribution that wraps the
 * chemistry call in the reference: SUBROUTINE stem_kpp, /root/reference/src/str.f90
 *   snapshot before kpp_driver   str.f90:5920-5966  (fs, sap, smp, sion1o)
 *   redistribution of ff along   str.f90:5976-6100  (den, x0, ix/c0, moves, vc)
 *   the dry-aerosol axis
 *   exchange of sl1/sion1        str.f90:6102-6134
 * Loop nests, loop directions and the order of every floating-point accumulation
 * follow the Fortran statement by statement (1-based indices kept in comments).
 *
 * PARITY STATUS: "parity unpinned" - the reference has no tests or fixtures for this
 * path and cannot be compiled here (no Fortran compiler); pinned by the property tests
 * in tests/test_bins_oracle.py (particle number per water bin conserved, dissolved mass
 * conserved across bins, identity for zero mass change, growth/shrink direction).
 *
 * Array layout (C row-major == Fortran column-major of the reference):
 *   ff    [ncell][nka][nkt]   = ff(nkt,nka,k)        particles cm^-3
 *   sion1 [ncell][nkc][j6]    = sion1(j6,nkc,k)      j6 = 55
 *   sl1   [ncell][nkc][j2]    = sl1(j2,nkc,k)        j2 = 121
 *   cw,cm,sap,smp [ncell][nkc]; sion1o [ncell][nkc][9]; rq [nka][nkt] = rq(nkt,nka)
 *   kw[nka], ka: 1-based limits exactly as in COMMON /blck06/.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NKC 4
#define LSP 9
#define J2 121
#define J6 55

/* lj2 of str.f90:5876 (1-based ion indices H+, NH4+, SO4=, HCO3-, NO3-, Cl-, HSO4-, Na+, CH3SO3-) */
static const int lj2[LSP] = {1, 2, 8, 9, 13, 14, 19, 20, 30};

typedef struct {
  int nka, nkt, ka, nkc_l, ial_first, reserved;
  const int *kw;
  const double *en;
  const double *rq;
} bins_grid;

static void ia_range(const bins_grid *g, int kc, int *ial, int *iau) /* 1-based, inclusive */
{
  if (kc == 1 || kc == 3) { *ial = g->ial_first; *iau = g->ka; }       /* str.f90:5924-5931 */
  else { *ial = g->ka + 1; *iau = g->nka - 1; }                        /* str.f90:5932-5935 */
}

static void jt_range(const bins_grid *g, int kc, int ia, int *jtl, int *jtu)
{
  if (kc == 1 || kc == 2) { *jtl = 1; *jtu = g->kw[ia - 1]; }          /* str.f90:5940-5942 */
  else { *jtl = g->kw[ia - 1] + 1; *jtu = g->nkt; }                    /* str.f90:5943-5946 */
}

/* str.f90:5920-5966 for one layer */
void bins_oracle_snapshot(const bins_grid *g, int64_t ncell, const double *ff, const double *cm,
                          const double *sion1, double *sap, double *smp, double *sion1o)
{
  const int nka = g->nka, nkt = g->nkt;
  for (int64_t c = 0; c < ncell; ++c) {
    const double *f = ff + (size_t)c * nka * nkt;
    for (int kc = 1; kc <= NKC; ++kc) {
      sap[c * NKC + kc - 1] = 0.0;                                     /* str.f90:5916-5917 */
      smp[c * NKC + kc - 1] = 0.0;
    }
    for (int kc = 1; kc <= g->nkc_l; ++kc) {
      if (cm[c * NKC + kc - 1] == 0.0) continue;                       /* str.f90:5922 */
      int ial, iau;
      ia_range(g, kc, &ial, &iau);
      double sapk = 0.0, smpk = 0.0;
      for (int ia = ial; ia <= iau; ++ia) {
        double fs = 0.0;
        int jtl, jtu;
        jt_range(g, kc, ia, &jtl, &jtu);
        for (int jt = jtl; jt <= jtu; ++jt) {
          const double v = f[(ia - 1) * nkt + (jt - 1)];
          fs = fs + v * g->en[ia - 1];                                 /* str.f90:5948 */
          sapk = sapk + v;                                             /* str.f90:5949 */
        }
        smpk = smpk + fs;                                              /* str.f90:5951 */
      }
      sap[c * NKC + kc - 1] = sapk;
      smp[c * NKC + kc - 1] = smpk;
      for (int l = 0; l < LSP; ++l)                                    /* str.f90:5959-5962 */
        sion1o[(c * NKC + kc - 1) * LSP + l] = sion1[(c * NKC + kc - 1) * J6 + lj2[l] - 1];
    }
  }
}

/* str.f90:5976-6134 for every layer; nwarn[c] counts the "aerosol growth" messages
 * (x0 <= 0, str.f90:6027) the reference writes to jpfunout. */
void bins_oracle_redistribute(const bins_grid *g, int64_t ncell, double *ff, const double *cm,
                              const double *cw, const double *sap, const double *smp,
                              const double *sion1o, double *sion1, double *sl1, int32_t *nwarn)
{
  const int nka = g->nka, nkt = g->nkt, ka = g->ka;
  const double pi = 3.1415926535897932;                                /* constants.f90 */
  const double fpi = 4.0 / 3.0 * pi;                                   /* str.f90:5838 */
  const float em6 = 1.e-06f;                                           /* default-REAL literal, str.f90:5983 */
  for (int64_t c = 0; c < ncell; ++c) {
    double *f = ff + (size_t)c * nka * nkt;
    double vc[NKC][NKC];                                               /* vc(tix,kc,k) */
    memset(vc, 0, sizeof vc);
    int warn = 0;
    for (int kc = 1; kc <= g->nkc_l; ++kc) {
      if (cm[c * NKC + kc - 1] == 0.0) continue;                       /* str.f90:5978 */
      const double sapk = sap[c * NKC + kc - 1], smpk = smp[c * NKC + kc - 1];
      if (!(sapk > 1.e-6)) continue;                                   /* str.f90:5979 */
      double ds[LSP];
      for (int l = 0; l < LSP; ++l)
        ds[l] = (sion1[(c * NKC + kc - 1) * J6 + lj2[l] - 1] - sion1o[(c * NKC + kc - 1) * LSP + l])
                * (double)em6 / sapk;                                  /* str.f90:5983 */
      const double den = (ds[0] * 1. + ds[1] * 18. + ds[2] * 96. + ds[3] * 44. + ds[4] * 62.
                          + ds[5] * 35.5 + ds[6] * 97. + ds[7] * 23. + ds[8] * 95.) * 1000.;  /* 5989-5993 */
      int ial, iau;
      ia_range(g, kc, &ial, &iau);
      int istart = ial, iend = iau, iinkr = 1;
      if (den >= 0.0) { istart = iau; iend = ial; iinkr = -1; }        /* str.f90:6016-6020 */
      for (int ia = istart; iinkr > 0 ? ia <= iend : ia >= iend; ia += iinkr) {
        const double x0 = g->en[ia - 1] + den * g->en[ia - 1] / smpk * sapk;   /* str.f90:6023-6026 */
        if (!(den > 0.0) && x0 <= 0.0) ++warn;
        int ix = 0;
        double c0 = 0.0;
        for (int iia = 1; iia <= nka - 1; ++iia) {                     /* str.f90:6030-6036 */
          if (g->en[iia - 1] <= x0 && g->en[iia] > x0) {
            ix = iia;
            c0 = (g->en[iia] - x0) / (g->en[iia] - g->en[iia - 1]);
            break;
          }
        }
        if (ix == 0) {                                                 /* str.f90:6037-6043 */
          if (g->en[0] > x0) { ix = 1; c0 = 1.0; }
          else { ix = nka - 1; c0 = 0.0; }
        }
        int jtl, jtu;
        jt_range(g, kc, ia, &jtl, &jtu);
        for (int jt = jtl; jt <= jtu; ++jt) {
          double *src = &f[(ia - 1) * nkt + (jt - 1)];
          if (*src > 0.0) {                                            /* str.f90:6055-6060 */
            const double x1 = *src;
            *src = 0.0;
            f[(ix - 1) * nkt + (jt - 1)] = f[(ix - 1) * nkt + (jt - 1)] + x1 * c0;
            f[ix * nkt + (jt - 1)] = f[ix * nkt + (jt - 1)] + x1 * (1.0 - c0);
            int tix, tixp;                                             /* str.f90:6061-6088 */
            if (ix > ka) tix = (jt > g->kw[ix - 1]) ? 4 : 2;
            else tix = (jt > g->kw[ix - 1]) ? 3 : 1;
            if (ix + 1 > ka) tixp = (jt > g->kw[ix]) ? 4 : 2;
            else tixp = (jt > g->kw[ix]) ? 3 : 1;
            const double r = g->rq[(ia - 1) * nkt + (jt - 1)];
            const double r3 = r * r * r;                               /* rq(jt,ia)**3 */
            if (tix != kc) vc[kc - 1][tix - 1] = vc[kc - 1][tix - 1] + x1 * c0 * fpi * r3;
            if (tixp != kc) vc[kc - 1][tixp - 1] = vc[kc - 1][tixp - 1] + x1 * (1.0 - c0) * fpi * r3;
          }
        }
      }
    }
    /* str.f90:6102-6134: move the dissolved species with the transferred volume */
    for (int kc = 1; kc <= g->nkc_l; ++kc) {
      for (int kkc = 1; kkc <= g->nkc_l; ++kkc) {
        if (kkc == kc) continue;
        if (vc[kc - 1][kkc - 1] == 0.0) continue;
        const double cwf = cw[c * NKC + kc - 1];
        if (cwf > 0.0) {
          const double vol_ch = vc[kc - 1][kkc - 1] * 1.e-12;
          const double xfact = 1.0 - (cwf - vol_ch) / cwf;
          double *slf = sl1 + (size_t)(c * NKC + kc - 1) * J2, *slt = sl1 + (size_t)(c * NKC + kkc - 1) * J2;
          for (int l = 0; l < J2; ++l) {
            const double xch = slf[l] * xfact;
            slf[l] = slf[l] - xch;
            slt[l] = slt[l] + xch;
          }
          double *sif = sion1 + (size_t)(c * NKC + kc - 1) * J6, *sit = sion1 + (size_t)(c * NKC + kkc - 1) * J6;
          for (int l = 0; l < J6; ++l) {
            const double xch = sif[l] * xfact;
            sif[l] = sif[l] - xch;
            sit[l] = sit[l] + xch;
          }
        }
      }
    }
    if (nwarn) nwarn[c] = warn;
  }
}
