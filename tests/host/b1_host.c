/* A minimal "Fortran host" for boundary B1 (test infrastructure): DEFINES the COMMON blocks /GDATA_g/, /GDATA_a/,
 * /GDATA_t/ with the reference's layout (gas_Global.h:28-58), fills them the way gas_drive / aer_drive / tot_drive
 * do before CALL INTEGRATE_x (gas.f:173, aer.f:217, tot.f:604) and calls the shims of libmistra_kpp_f77.so.
 * Built as a shared object so that the Python tests and bench.py can drive it through ctypes. */
#include "../../include/mistra_kpp_f77.h"

#include <string.h>
#include <time.h>

struct mistra_gdata_g gdata_g_;
struct mistra_gdata_a gdata_a_;
struct mistra_gdata_t gdata_t_;

static const int NVAR[3] = {102, 257, 417}, NFIX[3] = {3, 5, 7}, NREACT[3] = {331, 979, 1627};

static void parts(int mech, double **C, double **RC, double **stepmin, double **atol, double **rtol)
{
  if (mech == 0) { *C = gdata_g_.C; *RC = gdata_g_.RCONST; *stepmin = &gdata_g_.STEPMIN; *atol = gdata_g_.ATOL; *rtol = gdata_g_.RTOL; }
  else if (mech == 1) { *C = gdata_a_.C; *RC = gdata_a_.RCONST; *stepmin = &gdata_a_.STEPMIN; *atol = gdata_a_.ATOL; *rtol = gdata_a_.RTOL; }
  else { *C = gdata_t_.C; *RC = gdata_t_.RCONST; *stepmin = &gdata_t_.STEPMIN; *atol = gdata_t_.ATOL; *rtol = gdata_t_.RTOL; }
}

static void call(int mech, double *tin, double *tout)
{
  if (mech == 0) integrate_g_(tin, tout);
  else if (mech == 1) integrate_a_(tin, tout);
  else integrate_t_(tin, tout);
}

/* one cell through INTEGRATE_x: returns TIN after the call (= Texit), STEPMIN (= Hexit), ATOL(1), RTOL(1) */
int b1_integrate(int mech, const double *var, const double *fix, const double *rconst, double *var_out, double tin,
                 double tout, double *texit, double *stepmin, double *tol)
{
  double *C, *RC, *sm, *atol, *rtol;
  if (mech < 0 || mech > 2) return -1;
  parts(mech, &C, &RC, &sm, &atol, &rtol);
  memcpy(C, var, sizeof(double) * NVAR[mech]);
  memcpy(C + NVAR[mech], fix, sizeof(double) * NFIX[mech]);
  memcpy(RC, rconst, sizeof(double) * NREACT[mech]);
  call(mech, &tin, &tout);
  memcpy(var_out, C, sizeof(double) * NVAR[mech]);
  *texit = tin;
  *stepmin = *sm;
  tol[0] = atol[0];
  tol[1] = rtol[NVAR[mech] - 1];
  return 0;
}

/* the box-model pattern (kpp.f90:4296-4299): `ncalls` consecutive one-cell calls, VAR carried in the COMMON block,
 * restarted from `var` every `restart` calls; returns microseconds per call */
double b1_latency_us(int mech, const double *var, const double *fix, const double *rconst, int ncalls, int restart)
{
  double *C, *RC, *sm, *atol, *rtol;
  struct timespec t0, t1;
  int i;
  if (mech < 0 || mech > 2 || ncalls < 1) return -1.0;
  parts(mech, &C, &RC, &sm, &atol, &rtol);
  memcpy(C + NVAR[mech], fix, sizeof(double) * NFIX[mech]);
  memcpy(RC, rconst, sizeof(double) * NREACT[mech]);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (i = 0; i < ncalls; i++) {
    double tin = 0.0, tout = 10.0;
    if (restart > 0 && i % restart == 0) memcpy(C, var, sizeof(double) * NVAR[mech]);
    call(mech, &tin, &tout);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return ((t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3) / ncalls;
}
