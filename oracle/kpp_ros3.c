/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h).
 *
 * Hand restatement, statement by statement, of the reference's Rosenbrock
 * driver as configured by INTEGRATE_x.  Line numbers cite
 * /root/reference/src/gas.f; aer.f (1408, 1810, ...) and tot.f (2812, 3214, ...)
 * hold bodies identical modulo sizes (diffed).
 *
 * Compile WITHOUT floating-point contraction (-ffp-contract=off): the Fortran
 * reference built for generic x86-64 has no fused multiply-add.
 */
#include "kpp_oracle.h"
#include "kpp_oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

extern const kpp_mech_t kpp_mech_g_f32, kpp_mech_g_f64, kpp_mech_a_f32, kpp_mech_a_f64,
    kpp_mech_t_f32, kpp_mech_t_f64;

static const kpp_mech_t *get_mech(int mech, int f32)
{
  switch (mech) {
  case 0: return f32 ? &kpp_mech_g_f32 : &kpp_mech_g_f64;
  case 1: return f32 ? &kpp_mech_a_f32 : &kpp_mech_a_f64;
  case 2: return f32 ? &kpp_mech_t_f32 : &kpp_mech_t_f64;
  }
  return 0;
}

#define FMIN(a, b) ((a) < (b) ? (a) : (b))
#define FMAX(a, b) ((a) > (b) ? (a) : (b))

/* COMMON /Statistics/ (gas.f:913-915) - per call here, not shared */
typedef struct { int Nfun, Njac, Nstp, Nacc, Nrej, Ndec, Nsol, Nsng; } stats_t;

typedef struct {
  const kpp_mech_t *m;
  const double *FIX, *RCONST;    /* frozen inputs of the integration */
  stats_t st;
  /* locals of RosenbrockIntegrator_x (gas.f:1154-1160) */
  double *Ynew, *Fcn0, *Fcn, *K, *dFdT, *Jac0, *Ghimj, *Yerr, *W;
} work_t;

/* KppDecomp_x, gas.f:6142-6177: index-driven row-wise sparse LU, work row W. */
static int kpp_decomp(const kpp_mech_t *m, double *JVS, double *W)
{
  const int *LU_ICOL = m->lu_icol, *LU_CROW = m->lu_crow, *LU_DIAG = m->lu_diag;
  int k, kk, j, jj;
  double a;
  for (k = 0; k < m->nvar; k++) {
    if (JVS[LU_DIAG[k]] == 0.) return k + 1;                    /* gas.f:6156 */
    for (kk = LU_CROW[k]; kk < LU_CROW[k + 1]; kk++) W[LU_ICOL[kk]] = JVS[kk];
    for (kk = LU_CROW[k]; kk < LU_DIAG[k]; kk++) {
      j = LU_ICOL[kk];
      a = -W[j] / JVS[LU_DIAG[j]];                              /* gas.f:6164 */
      W[j] = -a;
      for (jj = LU_DIAG[j] + 1; jj < LU_CROW[j + 1]; jj++)
        W[LU_ICOL[jj]] = W[LU_ICOL[jj]] + a * JVS[jj];          /* gas.f:6167 */
    }
    for (kk = LU_CROW[k]; kk < LU_CROW[k + 1]; kk++) JVS[kk] = W[LU_ICOL[kk]];
  }
  return 0;
}

/* WAXPY_x, gas.f:6641-6673 (unroll-by-4 does not change per-element results) */
static void waxpy(int n, double alpha, const double *x, double *y)
{
  int i;
  if (alpha == 0.0) return;
  for (i = 0; i < n; i++) y[i] = y[i] + alpha * x[i];
}

/* FunTemplate_x / JacTemplate_x, gas.f:1946-2003: TIME is set but Fun ignores it. */
static void fun_template(work_t *w, double T, const double *Y, double *Ydot)
{
  (void)T;
  w->m->fun(Y, w->FIX, w->RCONST, Ydot);
  w->st.Nfun++;
}
static void jac_template(work_t *w, double T, const double *Y, double *Jcb)
{
  (void)T;
  w->m->jac(Y, w->FIX, w->RCONST, Jcb);
  w->st.Njac++;
}

/* ros_ErrorNorm_x, gas.f:1341-1372 (scalar tolerances: VectorTol = .FALSE.) */
static double ros_error_norm(int n, const double *Y, const double *Ynew, const double *Yerr,
                             double AbsTol, double RelTol)
{
  double Err = 0.0, Scale, Ymax, q;
  int i;
  for (i = 0; i < n; i++) {
    Ymax = FMAX(fabs(Y[i]), fabs(Ynew[i]));
    Scale = AbsTol + RelTol * Ymax;
    q = Yerr[i] / Scale;
    Err = Err + q * q;
  }
  return sqrt(Err / n);
}

/* ros_FunTimeDerivative_x, gas.f:1375-1400 */
static void ros_fun_time_derivative(work_t *w, double T, double Roundoff, const double *Y,
                                    const double *Fcn0, double *dFdT)
{
  const double DeltaMin = 1.0e-6;
  int i, n = w->m->nvar;
  double Delta = sqrt(Roundoff) * FMAX(DeltaMin, fabs(T));
  fun_template(w, T + Delta, Y, dFdT);
  waxpy(n, -1.0, Fcn0, dFdT);
  for (i = 0; i < n; i++) dFdT[i] = (1.0 / Delta) * dFdT[i];
}

/* ros_PrepareMatrix_x, gas.f:1404-1470 */
static int ros_prepare_matrix(work_t *w, double *H, int Direction, double gam)
{
  const kpp_mech_t *m = w->m;
  int i, ising, Nconsecutive = 0;
  double ghinv;
  for (;;) {
    for (i = 0; i < m->lu_nonzero; i++) w->Ghimj[i] = -w->Jac0[i];     /* gas.f:1445 */
    ghinv = 1.0 / (Direction * *H * gam);
    for (i = 0; i < m->nvar; i++) w->Ghimj[m->lu_diag[i]] = w->Ghimj[m->lu_diag[i]] + ghinv;
    ising = kpp_decomp(m, w->Ghimj, w->W);
    w->st.Ndec++;                                                       /* gas.f:1917 */
    if (ising == 0) return 0;
    w->st.Nsng++;
    Nconsecutive++;
    if (Nconsecutive <= 5) *H = *H * 0.5;                               /* gas.f:1459 */
    else return 1;
  }
}

/* RosenbrockIntegrator_x, gas.f:1112-1337, with the Ros3 tableau of gas.f:1570-1628 */
static int ros_integrator(work_t *w, double *Y, double Tstart, double Tend, double *Tout,
                          double AbsTol, double RelTol, int Autonomous, int Max_no_steps,
                          double Roundoff, double Hmin, double Hmax, double Hstart,
                          double *Hexit, double FacMin, double FacMax, double FacRej,
                          double FacSafe)
{
  enum { ros_S = 3 };
  static const double ros_A[3] = { 1.0, 1.0, 0.0 };
  static const double ros_C[3] = { -0.10156171083877702091975600115545e+01,
                                   0.40759956452537699824805835358067e+01,
                                   0.92076794298330791242156818474003e+01 };
  static const int ros_NewF[3] = { 1, 1, 0 };
  static const double ros_M[3] = { 0.1e+01, 0.61697947043828245592553615689730e+01,
                                   -0.42772256543218573326238373806514e+00 };
  static const double ros_E[3] = { 0.5e+00, -0.29079558716805469821718236208017e+01,
                                   0.22354069897811569627360909276199e+00 };
  static const double ros_Alpha[3] = { 0.0, 0.43586652150845899941601945119356e+00,
                                       0.43586652150845899941601945119356e+00 };
  static const double ros_Gamma[3] = { 0.43586652150845899941601945119356e+00,
                                       0.24291996454816804366592249683314e+00,
                                       0.21851380027664058511513169485832e+01 };
  const double ros_ELO = 3.0, DeltaMin = 1.0e-5;
  const kpp_mech_t *m = w->m;
  const int N = m->nvar;
  double *Ynew = w->Ynew, *Fcn0 = w->Fcn0, *Fcn = w->Fcn, *K = w->K, *dFdT = w->dFdT,
         *Yerr = w->Yerr;
  double T, H, Hnew, HC, HG, Fac, Tau, Err;
  int Direction, ioffset, i, j, istage, RejectLastH, RejectMoreH;

  T = Tstart;                                                          /* gas.f:1180 */
  *Hexit = 0.0;
  H = FMIN(Hstart, Hmax);
  if (fabs(H) <= 10.0 * Roundoff) H = DeltaMin;
  Direction = (Tend >= Tstart) ? +1 : -1;
  RejectLastH = 0;
  RejectMoreH = 0;

  while (fabs(Tend - T) >= Roundoff) {                                  /* gas.f:1202 */
    if (w->st.Nstp > Max_no_steps) { *Tout = T; return -6; }            /* gas.f:1204 */
    if (((T + 0.1 * H) == T) || (H <= Roundoff)) { *Tout = T; return -7; } /* gas.f:1208 */
    *Hexit = H;                                                         /* gas.f:1214 */
    H = FMIN(H, fabs(Tend - T));
    fun_template(w, T, Y, Fcn0);                                        /* gas.f:1218 */
    if (!Autonomous) ros_fun_time_derivative(w, T, Roundoff, Y, Fcn0, dFdT);
    jac_template(w, T, Y, w->Jac0);                                     /* gas.f:1227 */

    for (;;) {                                                          /* gas.f:1230 */
      if (ros_prepare_matrix(w, &H, Direction, ros_Gamma[0])) { *Tout = T; return -8; }
      for (istage = 1; istage <= ros_S; istage++) {
        ioffset = N * (istage - 1);
        if (istage == 1) {
          memcpy(Fcn, Fcn0, N * sizeof(double));
        } else if (ros_NewF[istage - 1]) {
          memcpy(Ynew, Y, N * sizeof(double));
          for (j = 1; j <= istage - 1; j++)
            waxpy(N, ros_A[(istage - 1) * (istage - 2) / 2 + j - 1], K + N * (j - 1), Ynew);
          Tau = T + ros_Alpha[istage - 1] * Direction * H;
          fun_template(w, Tau, Ynew, Fcn);
        }
        memcpy(K + ioffset, Fcn, N * sizeof(double));
        for (j = 1; j <= istage - 1; j++) {
          HC = ros_C[(istage - 1) * (istage - 2) / 2 + j - 1] / (Direction * H);
          waxpy(N, HC, K + N * (j - 1), K + ioffset);
        }
        if (!Autonomous && ros_Gamma[istage - 1] != 0.0) {
          HG = Direction * H * ros_Gamma[istage - 1];
          waxpy(N, HG, dFdT, K + ioffset);
        }
        m->solve(w->Ghimj, K + ioffset);                                /* gas.f:1274 */
        w->st.Nsol++;
      }
      memcpy(Ynew, Y, N * sizeof(double));                              /* gas.f:1281 */
      for (j = 1; j <= ros_S; j++) waxpy(N, ros_M[j - 1], K + N * (j - 1), Ynew);
      for (i = 0; i < N; i++) Yerr[i] = 0.0;
      for (j = 1; j <= ros_S; j++) waxpy(N, ros_E[j - 1], K + N * (j - 1), Yerr);
      Err = ros_error_norm(N, Y, Ynew, Yerr, AbsTol, RelTol);          /* gas.f:1294 */

      /* gas.f:1297.  Err==0 gives FacSafe/0 = +Inf -> FacMax (trap 2).  A NaN
       * Err is processor dependent in Fortran MIN/MAX; defined here (SURVEY §8a
       * trap 9) as "reject with Fac = FacMin". */
      if (Err != Err) Fac = FacMin;
      else Fac = FMIN(FacMax, FMAX(FacMin, FacSafe / pow(Err, 1.0 / ros_ELO)));
      Hnew = H * Fac;

      w->st.Nstp++;                                                     /* gas.f:1301 */
      if ((Err <= 1.0) || (H <= Hmin)) {
        w->st.Nacc++;
        memcpy(Y, Ynew, N * sizeof(double));
        T = T + Direction * H;
        Hnew = FMAX(Hmin, FMIN(Hnew, Hmax));
        if (RejectLastH) Hnew = FMIN(Hnew, H);
        RejectLastH = 0;
        RejectMoreH = 0;
        H = Hnew;
        break;
      } else {
        if (RejectMoreH) Hnew = H * FacRej;
        RejectMoreH = RejectLastH;
        RejectLastH = 1;
        H = Hnew;
        if (w->st.Nacc >= 1) w->st.Nrej++;
      }
    }
  }
  *Tout = T;
  return 1;                                                             /* gas.f:1333 */
}

/* Rosenbrock_x option decoding, gas.f:777-1108 (Ros3 only: IPAR(4)=2) */
static int rosenbrock(work_t *w, double *Y, double Tstart, double Tend, const kpp_oracle_opts *o,
                      double *Texit, double *Hexit)
{
  const double Roundoff = 2.220446049250313e-16;  /* epsilon(ONE), gas.f:972 */
  const double DeltaMin = 1.0e-5;
  double Hmin, Hmax, Hstart, FacMin, FacMax, FacRej, FacSafe;
  int Max_no_steps;
  memset(&w->st, 0, sizeof(w->st));
  *Texit = Tstart;
  *Hexit = 0.0;
  if (o->max_steps == 0) Max_no_steps = 100000;
  else if (o->max_steps > 0) Max_no_steps = o->max_steps;
  else return -1;
  if (o->hmin == 0.0) Hmin = 0.0; else if (o->hmin > 0.0) Hmin = o->hmin; else return -3;
  if (o->hmax == 0.0) Hmax = fabs(Tend - Tstart);
  else if (o->hmax > 0.0) Hmax = FMIN(fabs(o->hmax), fabs(Tend - Tstart)); else return -3;
  if (o->hstart == 0.0) Hstart = FMAX(Hmin, DeltaMin);
  else if (o->hstart > 0.0) Hstart = FMIN(fabs(o->hstart), fabs(Tend - Tstart)); else return -3;
  if (o->facmin == 0.0) FacMin = 0.2; else if (o->facmin > 0.0) FacMin = o->facmin; else return -4;
  if (o->facmax == 0.0) FacMax = 6.0; else if (o->facmax > 0.0) FacMax = o->facmax; else return -4;
  if (o->facrej == 0.0) FacRej = 0.1; else if (o->facrej > 0.0) FacRej = o->facrej; else return -4;
  if (o->facsafe == 0.0) FacSafe = 0.9; else if (o->facsafe > 0.0) FacSafe = o->facsafe; else return -4;
  if ((o->atol <= 0.0) || (o->rtol <= 10.0 * Roundoff) || (o->rtol >= 1.0)) return -5;
  return ros_integrator(w, Y, Tstart, Tend, Texit, o->atol, o->rtol, o->autonomous,
                        Max_no_steps, Roundoff, Hmin, Hmax, Hstart, Hexit, FacMin, FacMax,
                        FacRej, FacSafe);
}

static int work_alloc(work_t *w, const kpp_mech_t *m)
{
  size_t n = m->nvar, nz = m->lu_nonzero;
  double *p = (double *)calloc(n * 9 + nz * 2, sizeof(double));
  if (!p) return -1;
  w->m = m;
  w->Ynew = p; p += n;
  w->Fcn0 = p; p += n;
  w->Fcn = p; p += n;
  w->K = p; p += 3 * n;
  w->dFdT = p; p += n;
  w->Yerr = p; p += n;
  w->W = p; p += n;
  w->Jac0 = p; p += nz;
  w->Ghimj = p;
  return 0;
}

void kpp_oracle_default_opts(kpp_oracle_opts *o)
{
  memset(o, 0, sizeof(*o));
  o->rtol = 1.0e-3;      /* gas.f:745 */
  o->atol = 1.0e-25;     /* gas.f:746 */
  o->hstart = 1.0e-3;    /* gas.f:743 */
  o->f32_literals = 1;
}

int kpp_oracle_query(int mech, int *nvar, int *nfix, int *nreact, int *lu_nonzero)
{
  const kpp_mech_t *m = get_mech(mech, 1);
  if (!m) return -1;
  if (nvar) *nvar = m->nvar;
  if (nfix) *nfix = m->nfix;
  if (nreact) *nreact = m->nreact;
  if (lu_nonzero) *lu_nonzero = m->lu_nonzero;
  return 0;
}

const char *kpp_oracle_spc_name(int mech, int i)
{
  const kpp_mech_t *m = get_mech(mech, 1);
  if (!m || i < 0 || i >= m->nvar + m->nfix) return 0;
  return m->spc_names[i];
}

int kpp_oracle_tables(int mech, const int **icol, const int **crow, const int **diag)
{
  const kpp_mech_t *m = get_mech(mech, 1);
  if (!m) return -1;
  *icol = m->lu_icol; *crow = m->lu_crow; *diag = m->lu_diag;
  return 0;
}

int kpp_oracle_integrate(int mech, int64_t ncell, const double *rconst, const double *fix,
                         double *var, double t0, double t1, const kpp_oracle_opts *o,
                         int32_t *ierr, int32_t *stats, double *hexit, double *texit,
                         int nthreads)
{
  kpp_oracle_opts dflt;
  const kpp_mech_t *m;
  int fail = 0;
  if (!o) { kpp_oracle_default_opts(&dflt); o = &dflt; }
  m = get_mech(mech, o->f32_literals);
  if (!m) return -1;
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    work_t w;
    int64_t c;
    if (work_alloc(&w, m)) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      fail = 1;
    } else {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
      for (c = 0; c < ncell; c++) {
        double tx, hx;
        int ie;
        w.RCONST = rconst + c * m->nreact;
        w.FIX = fix + c * m->nfix;
        ie = rosenbrock(&w, var + c * m->nvar, t0, t1, o, &tx, &hx);
        if (ierr) ierr[c] = ie;
        if (hexit) hexit[c] = hx;
        if (texit) texit[c] = tx;
        if (stats) {
          int32_t *s = stats + 8 * c;
          s[0] = w.st.Nfun; s[1] = w.st.Njac; s[2] = w.st.Nstp; s[3] = w.st.Nacc;
          s[4] = w.st.Nrej; s[5] = w.st.Ndec; s[6] = w.st.Nsol; s[7] = w.st.Nsng;
        }
      }
      free(w.Ynew);
    }
  }
  return fail ? -2 : 0;
}

void kpp_oracle_fun(int mech, int f32, const double *V, const double *F, const double *RCT, double *Vdot)
{
  get_mech(mech, f32)->fun(V, F, RCT, Vdot);
}
void kpp_oracle_jac(int mech, int f32, const double *V, const double *F, const double *RCT, double *JVS)
{
  get_mech(mech, f32)->jac(V, F, RCT, JVS);
}
int kpp_oracle_decomp(int mech, double *JVS)
{
  const kpp_mech_t *m = get_mech(mech, 1);
  double *W = (double *)calloc(m->nvar, sizeof(double));
  int r = kpp_decomp(m, JVS, W);
  free(W);
  return r;
}
void kpp_oracle_solve(int mech, const double *JVS, double *X)
{
  get_mech(mech, 1)->solve(JVS, X);
}
