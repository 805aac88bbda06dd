/* mistra_difc.h - C ABI of SUBROUTINE difc and SUBROUTINE difp on the device (part of libmistra_kpp.so).
 *
 * First piece of "next" row N4 of the hot-path scope (the vertical operators on the
 * chemistry arrays): replaces SUBROUTINE difc (/root/reference/src/str.f90:3271-3445, called once per
 * time step from the main loop) for an ensemble of independent columns - the fully implicit turbulent
 * exchange of every chemical species (one tridiagonal system of nm - 1 levels per species and column,
 * solved by the reference's own forward / backward recurrences, all species of a column sharing the
 * matrix) followed by the explicit large-scale subsidence step.  Works in place on the arrays that
 * the chemistry kernels (mistra_kpp_integrate*, mistra_konc) use, so they can stay on the device
 * between the chemistry steps.
 *
 * Arrays (level index last in Fortran = slowest here, column slowest of all):
 *   atkh, w, am3 [ncol][n]    COMMON /cb42/, /cb45/, /blck01/   (level 1 of the reference = index 0)
 *   detw, deta   [n]          COMMON /cb41/ (the vertical grid, shared by the columns)
 *   field[i].s   [ncol][n][row]  IN/OUT; species 0 .. nproc-1 of each row are diffused, the others
 *                (bins kc > nkc_l of sl1 / sion1) and the levels 1 and n stay untouched:
 *                  s1(j1,n)        row = j1,      nproc = j1          (gas_common)
 *                  s3(j5,n)        row = j5,      nproc = j5
 *                  sl1(j2,nkc,n)   row = j2*nkc,  nproc = j2*nkc_l    COMMON /blck17/
 *                  sion1(j6,nkc,n) row = j6*nkc,  nproc = j6*nkc_l
 * Numerics: binary64, the reference's expressions and operation order, no FMA contraction, IEEE
 * division: every species is a sequential recurrence evaluated by one thread, so the results are
 * bit-identical to the reference order (tests compare with == against the oracle).
 * Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).  No CPU fallback. */
#ifndef MISTRA_DIFC_H
#define MISTRA_DIFC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MISTRA_DIFC_MAXFIELDS 8

typedef struct mistra_difc_field {
  double *s;
  int32_t row, nproc;
} mistra_difc_field;

typedef struct mistra_difc_args {
  int32_t n;            /* number of levels (global_params n); the systems span levels 2 .. n-1 */
  int32_t nfield;       /* <= MISTRA_DIFC_MAXFIELDS */
  double dt;
  const double *atkh, *w, *am3;
  const double *detw, *deta;
  mistra_difc_field field[MISTRA_DIFC_MAXFIELDS];
} mistra_difc_args;

/* HOST buffers (staged to the current device and back; synchronous). */
int mistra_difc(int64_t ncol, const mistra_difc_args *a, void *stream);
/* Every pointer is a DEVICE pointer on the current device; asynchronous on `stream`.  The
 * per-column matrix coefficients live in a library-owned scratch buffer (4 n doubles per column). */
int mistra_difc_device(int64_t ncol, const mistra_difc_args *d_a, void *stream);

/* SUBROUTINE difp (/root/reference/src/str.f90:3137-3265, called beside difc: str.f90:353): the same
 * implicit exchange + subsidence for the 2-D particle spectrum, nka * nkt tridiagonal systems per
 * column sharing difc's matrix, on ff / rho (the reference divides every level 2..n by rho, solves,
 * multiplies back - so level n is rewritten as ff / rho * rho - and only then applies the subsidence),
 * followed by fsum(k) = sum of ff(:,:,k) for k = 2..n.
 *   ff [ncol][n][row] IN/OUT, row = nka * nkt  COMMON /cb52/     rho [ncol][n]  COMMON /cb53/
 *   fsum [ncol][n]    IN/OUT (level 1 untouched)
 * ff is bit-identical to the reference order; fsum is a sum of row non-negative terms formed per group
 * of 128 consecutive grid points (butterfly over the lanes, warps in order) and then over the groups in
 * order (the reference keeps one running sum): ~1e-14 relative (tests: 1e-13). */
typedef struct mistra_difp_args {
  int32_t n, row;
  double dt;
  const double *atkh, *w, *rho;
  const double *detw, *deta;
  double *ff, *fsum;
} mistra_difp_args;

int mistra_difp(int64_t ncol, const mistra_difp_args *a, void *stream);              /* HOST buffers */
int mistra_difp_device(int64_t ncol, const mistra_difp_args *d_a, void *stream);     /* DEVICE pointers */

/* launches of both routines */
int64_t mistra_difc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
