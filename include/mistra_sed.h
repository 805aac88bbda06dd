/* mistra_sed.h - C ABI of the gravitational-settling operators on the device (part of libmistra_kpp.so).
 *
 * Rest of "next" row N4 of the hot-path scope (the vertical operators on the chemistry and
 * particle arrays): replaces, for an ensemble of independent columns,
 *   SUBROUTINE sedp  /root/reference/src/str.f90:2257-2411  settling of the 2-D particle spectrum ff (called after
 *                                                           kon, str.f90:362), with the deposition diagnostics
 *   SUBROUTINE sedl  /root/reference/src/str.f90:2627-2787  settling of the aqueous species sl1, sion1 (str.f90:394)
 *   SUBROUTINE sedc  /root/reference/src/str.f90:2567-2596  dry deposition / emission of the gas species: the loop
 *                                                           over the species (the statements above it edit the
 *                                                           table vg once and stay host code)
 * including SUBROUTINE advsed0 / advsed1 (str.f90:5522-5691, upstream and Bott's fourth-order positive definite
 * flux form on the levels 1..nf) and FUNCTION vterm (str.f90:2793-2864).  They work in place on the arrays the
 * chemistry kernels (mistra_kpp_integrate*, mistra_bins_*, mistra_kon_*, mistra_difc / mistra_difp) use, so a
 * column can stay on the device from one chemistry step to the next.
 *
 * Arrays (level index last in Fortran = slowest here, column slowest of all; level 1 of the reference = index 0):
 *   detw, deta [n]           COMMON /cb41/ (the vertical grid, shared by the columns)
 *   t, p [ncol][n]           COMMON /cb53/
 *   ff [ncol][n][nka][nkt]   COMMON /cb52/   IN/OUT, levels 2 .. nf
 *   rq [nka][nkt], e [nkt]   COMMON /cb50/   kw [nka] COMMON /blck06/ (1-based class index as the reference holds it)
 *   vd [ncol][nka][nkt], vdm [ncol][nkc], vt [ncol][n][nkc]   COMMON /kpp_vt/ (vt(nkc,nf): levels above nf unused)
 *   rc [ncol][n][nkc]        COMMON /blck11/
 *   sl1 [ncol][n][nkc][j2], sion1 [ncol][n][nkc][j6]   COMMON /blck17/   IN/OUT, levels 1 .. nf-1, bins 1 .. nkc_l
 *   s1 [ncol][n][j1]         gas_common     IN/OUT, levels 1 and 2
 *   diag [ncol][4]           ajs (OUT), trdep, ds1, ds2 (IN/OUT)   COMMON /cb47/
 * Numerics: binary64, the reference's expressions and operation order, no FMA contraction, IEEE division (advsed1's
 * divisions by constants through an equivalent correctly rounded sequence, see mistra_sed_divc_selftest); every
 * profile is a sequential recurrence in the level index evaluated by one thread, the diagnostics are summed over the
 * classes in the reference's order: results are bit-identical to the reference order wherever no exp / log enters
 * (advsed0 / advsed1 themselves, vterm below 10 um radius, hence sedl / sedp of all but large drops); vterm of large
 * drops and sedc carry CUDA's exp / log (1 ulp) - tests compare those at 1e-12.
 * Two statements of the reference's behaviour are kept on purpose: sedp's local x0 is only assigned for classes
 * that hold particles (column sum of ff * detw > 1e-6), so an empty class books the deposition of the class before
 * it once more (str.f90:2352, 2397; x0 before the first class is taken as 0 where Fortran leaves it undefined);
 * and ff(:,:,nf) is overwritten by level nf-1 for settled classes only.
 * The sub-step loops (do while dt0 > 0.1, str.f90:2355, 2708) end after at most MISTRA_SED_MAXSUB passes on the
 * device, so that inconsistent inputs (a zero settling velocity at level 2 in sedl) cannot hang it.
 * Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).  No CPU fallback. */
#ifndef MISTRA_SED_H
#define MISTRA_SED_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MISTRA_SED_MAXSUB 100000

typedef struct mistra_sedp_args {
  int32_t n, nf, nka, nkt;      /* levels, levels with microphysics (global_params nf), spectrum grid */
  double dt;
  const double *detw, *deta, *t, *p;
  const double *rq, *e;
  const int32_t *kw;
  const double *vd;
  double *ff, *diag;
} mistra_sedp_args;

int mistra_sedp(int64_t ncol, const mistra_sedp_args *a, void *stream);             /* HOST buffers, synchronous */
int mistra_sedp_device(int64_t ncol, const mistra_sedp_args *d_a, void *stream);    /* DEVICE pointers, asynchronous */

typedef struct mistra_sedl_args {
  int32_t n, nf, nkc, nkc_l, j2, j6;
  double dt;
  const double *detw, *deta, *t, *p;
  const double *rc, *vt, *vdm;
  double *sl1, *sion1;          /* either may be NULL */
} mistra_sedl_args;

int mistra_sedl(int64_t ncol, const mistra_sedl_args *a, void *stream);
int mistra_sedl_device(int64_t ncol, const mistra_sedl_args *d_a, void *stream);

typedef struct mistra_sedc_args {
  int32_t n, j1;
  double dt;
  const double *detw, *deta;
  const double *vg, *es1;       /* [j1]: deposition velocities as sedc's preamble leaves them, emission rates */
  double *s1;
} mistra_sedc_args;

int mistra_sedc(int64_t ncol, const mistra_sedc_args *a, void *stream);
int mistra_sedc_device(int64_t ncol, const mistra_sedc_args *d_a, void *stream);

/* Diagnostic: advsed1's divisions by 24, 48, 1920, 384, 768, 3840 are evaluated on the device as multiply + two FMAs
 * (correctly rounded by Markstein's theorem inside an exponent window, the IEEE division outside it).  This compares
 * that sequence with the IEEE division bit for bit on n pseudo-random and structured arguments (all exponents incl.
 * denormals / Inf / NaN, exactly divisible significands, significands of all ones) and returns the number of differing
 * results in *mismatches (expected 0).  Synchronous. */
int mistra_sed_divc_selftest(int64_t n, uint64_t seed, int64_t *mismatches);

/* launches of the three routines */
int64_t mistra_sed_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
