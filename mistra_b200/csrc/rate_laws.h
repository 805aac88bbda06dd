// Rate-law functions called by Update_RCONST_x — host-side restatement of
// /root/reference/src/kpp.f90:7127-8373 ("next" row N1).
//
// In v1 the rate constants are produced on the host exactly as the reference
// does (they are a frozen input of the integration); this header is written
// host/device-neutral so that the N1 device port can include it unchanged.
//
// Literal convention: the reference writes many constants as default-REAL
// literals (300., 8.314, 0.21, 4.60138, 10**(-6.16) ...).  RL(x) reproduces the
// value the reference's preferred flags give (binary32 promoted to binary64)
// when cx->f32 != 0, and the -r8 value otherwise (SURVEY §8a traps 1, 11).
#ifndef MISTRA_RATE_LAWS_H
#define MISTRA_RATE_LAWS_H

#include <math.h>

#if defined(__CUDACC__)
#define RL_HD __host__ __device__ static inline
#else
#define RL_HD static inline
#endif

typedef struct rate_ctx_s {
  // COMMON /cb_1/ aircc,te,h2oppm,pk  (kpp.f90:4315-4321)
  double aircc, te, h2oppm, pk;
  // COMMON /kpp_rate_x/ scalars (gas_Global.h:81-84, tot_Global.h:79-93)
  double conv1, xhal, xiod, xhet1, xhet2;
  double xliq[4], cvv[4];
  const double *ph_rat;   // [47]           COMMON /ph_r_x/
  const double *C;        // [NSPEC]        start-of-step concentrations (VAR | FIX)
  const double *FIX;      // = C + NVAR
  const double *yhenry;   // [NSPEC]
  const double *yxkmt;    // [nkc][NSPEC]   (Fortran yxkmt(NSPEC,nkc))
  const double *ykef;     // [nkc][NSPEC]
  const double *ykeb;     // [nkc][NSPEC]
  const double *yxkmtd;   // [2][NSPEC]
  const double *yxeq;     // [NSPEC]
  const double *ycw;      // [nkc]
  const double *ycwd;     // [2]
  int nspec;
  int f32;                // 1: default-REAL literals are binary32
  // 0-based species indices the heterogeneous-rate functions need (-1: absent)
  int i_HNO3, i_N2O5, i_NH3, i_H2SO4, i_ClNO3, i_BrNO3;
  int i_HNO3l[2], i_NO3ml[2], i_Clml[4], i_Brml[4];
  int if_H2Ol[4];         // index into FIX
} rate_ctx;

#define RL(x) (cx->f32 ? (double)x##f : (double)x)

// kpp.f90:7127  farr=a*exp(b/te)
RL_HD double farr(const rate_ctx *cx, double a, int b) { return a * exp(b / cx->te); }

// kpp.f90:7149  farr_sp=a*((te/b)**c)*exp(d/te)
RL_HD double farr_sp(const rate_ctx *cx, double a, int b, double c, int d)
{
  return a * pow(cx->te / b, c) * exp(d / cx->te);
}

// common tail of the Troe expressions, kpp.f90:7186
RL_HD double rl_troe(double a0, double b0, double x2)
{
  double l = log10(a0 / b0);
  return (a0 / (1 + a0 / b0)) * pow(x2, 1 / (1 + l * l));
}

// kpp.f90:7171 (300. is exact in binary32)
RL_HD double atk_3(const rate_ctx *cx, double a1, double a2, double b1, double b2, double fc)
{
  double a0 = a1 * cx->aircc * pow(cx->te / 300., a2);
  double b0 = b1 * pow(cx->te / 300., b2);
  return rl_troe(a0, b0, fc);
}

// kpp.f90:7304
RL_HD double atk_3f(const rate_ctx *cx, double a1, double a2, double b1, double b2, double fc)
{
  double a0 = a1 * cx->aircc * pow(cx->te / 298., a2);
  double b0 = b1 * pow(cx->te / 298., b2);
  return rl_troe(a0, b0, fc);
}

// kpp.f90:7327  statement function func(a0,b0)=a0*exp(b0*tte), tte=1./te
RL_HD double shno3(const rate_ctx *cx, double a1, int b1, double a2, int b2, double a3, int b3)
{
  double tte = 1. / cx->te;
  double f1 = a1 * exp(b1 * tte), f2 = a2 * exp(b2 * tte), f3 = a3 * exp(b3 * tte);
  return f1 + (f3 * cx->aircc / (1 + f3 * cx->aircc / f2));
}

// kpp.f90:7355
RL_HD double fbck(const rate_ctx *cx, double a1, double a2, double b1, double b2, double fc,
                  double ak, double bk)
{
  double a0 = a1 * cx->aircc * pow(cx->te / 300.0, a2);
  double b0 = b1 * pow(cx->te / 300.0, b2);
  return rl_troe(a0, b0, fc) / (ak * exp(bk / cx->te));
}

// kpp.f90:7383 (fc fixed to 0.6d0)
RL_HD double fbckj(const rate_ctx *cx, double a1, double a2, double b1, double b2, double ak,
                   double bk)
{
  return fbck(cx, a1, a2, b1, b2, 0.6, ak, bk);
}

// kpp.f90:7411  (8.314/101325. is a default-REAL quotient)
RL_HD double fbck2(const rate_ctx *cx, double a1, double a2, double b1, double b2, double fc,
                   double ck)
{
  const double ak = 5.44e-9, bk = 14192.0;
  double a0 = a1 * cx->aircc * pow(cx->te / 300., a2);
  double b0 = b1 * pow(cx->te / 300., b2);
  double x1 = rl_troe(a0, b0, fc);
  // kpp.f90:7437  x1/(ak*exp(bk/te)*8.314/101325.*te/ck): evaluated left to right, so each default-REAL literal is
  // promoted to double when it meets the running product - they are never divided by each other in single precision
  // (found by the independent evaluator tests/golden/make_rconst_reference.py: 1.7e-8 relative)
  if (ck != 0.0) return x1 / (ak * exp(bk / cx->te) * RL(8.314) / RL(101325.) * cx->te / ck);
  return 0.0;
}

// kpp.f90:7463
RL_HD double sp_17(const rate_ctx *cx, double a, double b) { return a * (1.0 + cx->aircc / b); }

// kpp.f90:7483
RL_HD double sp_23(const rate_ctx *cx, double a1, int b1, double a2, int b2, double a3, int b3)
{
  double tte = 1. / cx->te;
  return (a1 * exp(b1 * tte) + (a2 * cx->aircc) * exp(b2 * tte)) *
         (1 + ((a3 * cx->aircc * cx->h2oppm * 1.0e-6) * exp(b3 * tte)));
}

// kpp.f90:7540  fcn=10**(-6.16)*exp(-90.7d3/x2)*xmg*x1, x2=8.314*te
RL_HD double fcn(const rate_ctx *cx, double x1)
{
  double x2 = RL(8.314) * cx->te;
  double xmg = cx->pk / x2;
  double p10 = cx->f32 ? (double)powf(10.0f, -6.16f) : pow(10.0, -6.16);
  return p10 * exp(-90.7e3 / x2) * xmg * x1;
}

// kpp.f90:8351
RL_HD double dms_add(const rate_ctx *cx)
{
  double o2 = RL(0.21) * cx->aircc;
  double tte = 1. / cx->te;
  return 9.5e-39 * exp(5270. * tte) * o2 / (1.0 + 7.5e-29 * exp(5610. * tte) * o2);
}

// kpp.f90:7562
RL_HD double farr2(const rate_ctx *cx, double a0, int b0)
{
  return a0 * exp((double)b0 * (1.0 / cx->te - 3.3557e-3));
}

// kpp.f90:7687, 7708
RL_HD double dmin2(const rate_ctx *cx, double a) { (void)cx; return a < 1.e10 ? a : 1.e10; }
RL_HD double dmin3(const rate_ctx *cx, double a) { (void)cx; return a < 2.e10 ? a : 2.e10; }

// kpp.f90:7662
RL_HD double fliq_60(const rate_ctx *cx, double a1, int b1, double c, double d)
{
  if (d > 0.0) return a1 * exp((double)b1 * (1.0 / cx->te - 3.3557e-3)) * c / (c + 0.1 / d);
  return 0.0;
}

// kpp.f90:7757, 7779, 7801
RL_HD double flsc4(const rate_ctx *cx, double a, double b, double c)
{
  (void)cx;
  return c > 0.0 ? a * b * (c * c * c) : 0.0;
}
RL_HD double flsc5(const rate_ctx *cx, double a, double b, double c)
{
  (void)cx;
  return c > 0.0 ? a * (b * b) * ((c * c) * (c * c)) : 0.0;
}
RL_HD double flsc6(const rate_ctx *cx, double a, double b)
{
  (void)cx;
  return b > 1.e-15 ? a / b : 0.0;
}

// kpp.f90:7862
RL_HD double uplim(const rate_ctx *cx, double a, double b, double c, double d)
{
  (void)cx;
  if (d > 0.0) return a / (1.0 + b / 1.e10 * (c > 0.0 ? c : 0.0) * d);
  return 0.0;
}
// kpp.f90:7888
RL_HD double uparm(const rate_ctx *cx, double a0, int b0, double c, double d, double e)
{
  if (d > 0.0)
    return a0 * exp((double)b0 * (1.0 / cx->te - 3.3557e-3)) / (1.0 + c / 1.e10 * d * e);
  return 0.0;
}
// kpp.f90:7916
RL_HD double uplip(const rate_ctx *cx, double a, double b, double c)
{
  (void)cx;
  if (c > 0.0) return (a / (1.0 + a / 1.e10 * (b > 0.0 ? b : 0.0) * c) * (c * c));
  return 0.0;
}
// kpp.f90:7942
RL_HD double uparp(const rate_ctx *cx, double a0, int b0, double c, double d)
{
  if (d > 0.0) {
    double k = a0 * exp((double)b0 * (1.0 / cx->te - 3.3557e-3));
    return k / (1.0 + k / 1.e10 * c * d) * (d * d);
  }
  return 0.0;
}

#define RL_C(i) ((i) >= 0 ? cx->C[(i)] : 0.0)
#define RL_2D(arr, spc, k) (cx->arr[(k) * cx->nspec + (spc)])

// kpp.f90:7582  fhet_t(a0,b0,c0): bins a0=1..4
RL_HD double fhet_t(const rate_ctx *cx, int a0, int b0, int c0)
{
  int k = a0 - 1;
  double h2oa = cx->FIX[cx->if_H2Ol[k]];
  double hetT = h2oa + 5.0e2 * RL_C(cx->i_Clml[k]) + 3.0e5 * RL_C(cx->i_Brml[k]);
  double xbr = (b0 == 1) ? h2oa : (b0 == 2 ? 5.0e2 : 3.0e5);
  int sp = (c0 == 1) ? cx->i_N2O5 : (c0 == 2 ? cx->i_ClNO3 : cx->i_BrNO3);
  double xtr = RL_2D(yxkmt, sp, k);
  if (hetT > 0.0) return xtr * cx->ycw[k] * xbr / hetT;
  return 0.0;
}

// kpp.f90:8023 fhet_da / kpp.f90:8111 fhet_dt (same body for bins 1,2)
RL_HD double fhet_da(const rate_ctx *cx, double xliq, double xhet, int a0, int b0, int c0)
{
  int k = a0 - 1, sp;
  double h2oa, hetT, xbr, xtr, yw;
  (void)xliq;
  if (xhet == 0.) {
    sp = (c0 == 1) ? cx->i_N2O5 : (c0 == 2 ? cx->i_ClNO3 : cx->i_BrNO3);
    xtr = RL_2D(yxkmt, sp, k);
    h2oa = cx->FIX[cx->if_H2Ol[k]];
    hetT = h2oa + 5.0e2 * RL_C(cx->i_Clml[k]) + 3.0e5 * RL_C(cx->i_Brml[k]);
    yw = cx->ycw[k];
    if (cx->xhal == 0.) {
      if (c0 == 2 || c0 == 3) xtr = 0.;
      hetT = cx->FIX[cx->if_H2Ol[k]];
    }
  } else {
    // note the swapped species order of the dry branch (kpp.f90:8068-8070)
    sp = (c0 == 1) ? cx->i_N2O5 : (c0 == 2 ? cx->i_BrNO3 : cx->i_ClNO3);
    xtr = RL_2D(yxkmtd, sp, k);
    h2oa = RL(55.55) * cx->ycwd[k] * 1.e+3;
    hetT = h2oa + 5.0e2 * RL_C(cx->i_Clml[k]) + 3.0e5 * RL_C(cx->i_Brml[k]);
    yw = cx->ycwd[k];
    if (cx->xhal == 0.) hetT = RL(55.55) * cx->ycwd[k] * 1.e+3;
  }
  xbr = (b0 == 1) ? h2oa : (b0 == 2 ? 5.0e2 : 3.0e5);
  if (hetT > 0.0) return xtr * yw * xbr / hetT;
  return 0.0;
}
// kpp.f90:8111.  Differs from fhet_da only where xhal==0 in the wet branch:
// fhet_dt does not zero xtr for c0=2,3 (kpp.f90:8143-8146).
RL_HD double fhet_dt(const rate_ctx *cx, double xliq, double xhet, int a0, int b0, int c0)
{
  if (xhet == 0. && cx->xhal == 0. && (c0 == 2 || c0 == 3)) {
    int k = a0 - 1;
    int sp = (c0 == 2) ? cx->i_ClNO3 : cx->i_BrNO3;
    double h2oa = cx->FIX[cx->if_H2Ol[k]];
    double hetT = h2oa;
    double xbr = (b0 == 1) ? h2oa : (b0 == 2 ? 5.0e2 : 3.0e5);
    if (hetT > 0.0) return RL_2D(yxkmt, sp, k) * cx->ycw[k] * xbr / hetT;
    return 0.0;
  }
  return fhet_da(cx, xliq, xhet, a0, b0, c0);
}

// kpp.f90:8198  fdhetg(na,nb): gas mechanism, caq from HNO3lz*1.5d3
RL_HD double fdhetg(const rate_ctx *cx, int na, int nb)
{
  int k = na - 1;
  double xkt;
  if (nb == 1) {
    double x1 = RL_2D(yxkmtd, cx->i_HNO3, k) * cx->ycwd[k];
    double caq = ((RL_C(cx->i_HNO3l[k]) * 1.5e3) * 1.e-2) / (cx->yxeq[cx->i_HNO3] + 1.e-2);
    double x2 = 0.0;
    if (cx->C[cx->i_HNO3] != 0.0 && cx->yhenry[cx->i_HNO3] != 0.0)
      x2 = -RL_2D(yxkmtd, cx->i_HNO3, k) / (cx->C[cx->i_HNO3] * cx->yhenry[cx->i_HNO3]) * caq;
    xkt = (x1 + x2) > 0.0 ? (x1 + x2) : 0.0;
  } else {
    int sp = (nb == 2) ? cx->i_N2O5 : (nb == 3 ? cx->i_NH3 : cx->i_H2SO4);
    xkt = RL_2D(yxkmtd, sp, k) * cx->ycwd[k];
  }
  return xkt;
}

// kpp.f90:8269 fdheta / 8311 fdhett (identical bodies): caq from HNO3lz+NO3mlz
RL_HD double fdheta(const rate_ctx *cx, int na, int nb)
{
  int k = na - 1;
  double xkt;
  if (nb == 1) {
    double x1 = RL_2D(yxkmtd, cx->i_HNO3, k) * cx->ycwd[k];
    double caq = 0.0, x2 = 0.0;
    if ((cx->yxeq[cx->i_HNO3] + 1.e-2) != 0.0)
      caq = ((RL_C(cx->i_HNO3l[k]) + RL_C(cx->i_NO3ml[k])) * 1.e-2) / (cx->yxeq[cx->i_HNO3] + 1.e-2);
    if (cx->C[cx->i_HNO3] != 0.0 && cx->yhenry[cx->i_HNO3] != 0.0)
      x2 = -RL_2D(yxkmtd, cx->i_HNO3, k) / (cx->C[cx->i_HNO3] * cx->yhenry[cx->i_HNO3]) * caq;
    xkt = (x1 + x2) > 0.0 ? (x1 + x2) : 0.0;
  } else {
    int sp = (nb == 2) ? cx->i_N2O5 : (nb == 3 ? cx->i_NH3 : cx->i_H2SO4);
    xkt = RL_2D(yxkmtd, sp, k) * cx->ycwd[k];
  }
  return xkt;
}
RL_HD double fdhett(const rate_ctx *cx, int na, int nb) { return fdheta(cx, na, nb); }

// accessor macros used by the generated RCONST expressions
#define S_conv1 (cx->conv1)
#define S_xhal (cx->xhal)
#define S_xiod (cx->xiod)
#define S_xhet1 (cx->xhet1)
#define S_xhet2 (cx->xhet2)
#define S_xliq1 (cx->xliq[0])
#define S_xliq2 (cx->xliq[1])
#define S_xliq3 (cx->xliq[2])
#define S_xliq4 (cx->xliq[3])
#define S_cvv1 (cx->cvv[0])
#define S_cvv2 (cx->cvv[1])
#define S_cvv3 (cx->cvv[2])
#define S_cvv4 (cx->cvv[3])
#define PH_RAT(i) (cx->ph_rat[(i)])
#define YCW(i) (cx->ycw[(i)])
#define YHENRY(s) (cx->yhenry[(s)])
#define C_(s) (cx->C[(s)])
#define FIX_(s) (cx->FIX[(s)])
#define YXKMT(s, k) RL_2D(yxkmt, s, k)
#define YKEF(s, k) RL_2D(ykef, s, k)
#define YKEB(s, k) RL_2D(ykeb, s, k)

#endif
