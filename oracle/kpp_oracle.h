/* CPU ORACLE - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's KPP Rosenbrock path
 * (/root/reference/src/{gas,aer,tot}.f).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (mistra_b200/csrc, libmistra_kpp.so) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" - the reference ships no tests, golden
 * vectors or fixtures for this path and cannot be compiled
 * here (Fortran only, no Fortran compiler in the image).  The restatement is
 * generated statement by statement from the vendored KPP output and is pinned
 * by structure checks instead (tests/test_oracle_*.py): the unrolled KppSolve of
 * the reference equals CSR substitution over its own LU tables, Jac_SP equals
 * the analytic derivative of Fun, LU*x=b residuals, Ros3 order conditions and
 * a Radau cross-check.
 */
#ifndef KPP_ORACLE_H
#define KPP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same field meaning as RPAR/IPAR of Rosenbrock_x (gas.f:786-870): a zero
 * selects the reference default.  INTEGRATE_x (gas.f:739-746) sets
 * rtol=1e-3, atol=1e-25, hstart=1e-3, Ros3, non-autonomous, scalar tolerances. */
typedef struct {
  double rtol, atol;      /* RTOL(1), ATOL(1) (scalar tolerances, IPAR(2)=1)   */
  double hmin, hmax, hstart;            /* RPAR(1..3)                           */
  double facmin, facmax, facrej, facsafe; /* RPAR(4..7)                         */
  int32_t max_steps;      /* IPAR(3)                                            */
  int32_t autonomous;     /* IPAR(1)!=0 ; reference path uses 0                 */
  int32_t f32_literals;   /* 1: default-REAL literals are binary32 (preferred   */
                          /*    ifort/gfortran flags); 0: -r8 build             */
  int32_t reserved;
} kpp_oracle_opts;

void kpp_oracle_default_opts(kpp_oracle_opts *o);

/* mech: 0 gas, 1 aer, 2 tot */
int kpp_oracle_query(int mech, int *nvar, int *nfix, int *nreact, int *lu_nonzero);
const char *kpp_oracle_spc_name(int mech, int i);
int kpp_oracle_tables(int mech, const int **icol, const int **crow, const int **diag);

/* one INTEGRATE_x call per cell, cells [ncell][...] row-major; stats =
 * Nfun,Njac,Nstp,Nacc,Nrej,Ndec,Nsol,Nsng per cell; nthreads<=1: serial */
int kpp_oracle_integrate(int mech, int64_t ncell, const double *rconst, const double *fix,
                         double *var, double t0, double t1, const kpp_oracle_opts *o,
                         int32_t *ierr, int32_t *stats, double *hexit, double *texit,
                         int nthreads);

/* building blocks, for the structure tests */
void kpp_oracle_fun(int mech, int f32, const double *V, const double *F, const double *RCT, double *Vdot);
void kpp_oracle_jac(int mech, int f32, const double *V, const double *F, const double *RCT, double *JVS);
int kpp_oracle_decomp(int mech, double *JVS);
void kpp_oracle_solve(int mech, const double *JVS, double *X);

#ifdef __cplusplus
}
#endif
#endif
