/* mistra_kpp_f77.h - boundary B1: link-level replacements of the reference's
 * per-cell integrator entry points (libmistra_kpp_f77.a / .so).
 *
 * A Mistra build that drops INTEGRATE_g/_a/_t from gas.f/aer.f/tot.f (or links
 * this library first) gets these instead; everything else of the Fortran host is
 * unchanged.  Symbol names follow the gfortran/ifort convention (lower case,
 * trailing underscore, arguments by reference).  Each shim reads VAR/FIX/RCONST
 * from the mechanism's COMMON /GDATA_x/ (gas_Global.h:28-58: C(NSPEC) with
 * VAR = C(1:NVAR), FIX = C(NVAR+1:), RCONST(NREACT), TIME, DT, ATOL(NVAR),
 * RTOL(NVAR), STEPMIN, STEPMAX - all REAL*8, no padding), sets RTOL(:)=1e-3 and
 * ATOL(:)=1e-25 as INTEGRATE_x does (gas.f:745-746), integrates ONE cell on the
 * GPU through mistra_kpp_integrate, writes VAR back, and returns TIN = Texit,
 * STEPMIN = Hexit (gas.f:769-770).  On IERR < 0 it prints the reference's message
 * (gas.f:764-767) to stderr and returns normally with the partially advanced VAR.
 *
 * This is the latency path (one cell per call, ~tens of microseconds of launch
 * and copy overhead); the throughput path is B2, mistra_kpp_integrate with all
 * layers of a column (or many columns) in one call.
 */
#ifndef MISTRA_KPP_F77_H
#define MISTRA_KPP_F77_H
#ifdef __cplusplus
extern "C" {
#endif

/* replaces SUBROUTINE INTEGRATE_g(TIN,TOUT)  /root/reference/src/gas.f:710 */
void integrate_g_(double *tin, double *tout);
/* replaces SUBROUTINE INTEGRATE_a(TIN,TOUT)  /root/reference/src/aer.f:1408 */
void integrate_a_(double *tin, double *tout);
/* replaces SUBROUTINE INTEGRATE_t(TIN,TOUT)  /root/reference/src/tot.f:2812 */
void integrate_t_(double *tin, double *tout);

/* Optional: hand the shim the address of a mechanism's COMMON /GDATA_x/ image (mech = 0 gas, 1 aer, 2 tot) instead
 * of letting it look the symbol gdata_x_ up in the process (dlsym) at the first call. */
void mistra_kpp_f77_bind(int mech, void *gdata);

/* COMMON-block images the shims bind to (defined by the Fortran program). */
struct mistra_gdata_g { double C[105], RCONST[331], TIME, DT, ATOL[102], RTOL[102], STEPMIN, STEPMAX; };
struct mistra_gdata_a { double C[262], RCONST[979], TIME, DT, ATOL[257], RTOL[257], STEPMIN, STEPMAX; };
struct mistra_gdata_t { double C[424], RCONST[1627], TIME, DT, ATOL[417], RTOL[417], STEPMIN, STEPMAX; };

#ifdef __cplusplus
}
#endif
#endif
