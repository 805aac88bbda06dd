/* mistra_bins.h - C ABI of the 2-D aerosol/droplet bin redistribution that wraps the
 * chemistry step (part of libmistra_kpp.so).
 *
 * Replaces the loop nests of SUBROUTINE stem_kpp around its kpp_driver call
 * (/root/reference/src/str.f90:5797-6136):
 *
 *   mistra_bins_snapshot      str.f90:5916-5966  per (layer k, chem bin kc): number sap and
 *                                                dry mass smp of the bin's particles from
 *                                                ff(jt,ia,k), and the 9 mass-defining ions
 *                                                sion1o before chemistry
 *   [ kpp_driver -> mistra_kpp_integrate ]      str.f90:5971
 *   mistra_bins_redistribute  str.f90:5976-6134  mass change per particle den from the ion
 *                                                change, shift of every dry-mass class ia of
 *                                                the bin to its new dry mass (linear split
 *                                                over ix, ix+1), transferred volume vc, and
 *                                                the exchange of sl1/sion1 between bins
 *
 * Layers are independent, so `ncell` layers (of one or many columns) go in one call.
 * Arrays are C row-major = the reference's Fortran column-major arrays with the layer
 * index last:
 *   ff     [ncell][nka][nkt]  = ff(nkt,nka,k)   COMMON /cb52/     particles cm^-3, in/out
 *   cw, cm [ncell][4]         = cw(nkc,k), cm(nkc,k)  COMMON /blck12/
 *   sion1  [ncell][4][55]     = sion1(j6,nkc,k) COMMON /blck17/   in/out
 *   sl1    [ncell][4][121]    = sl1(j2,nkc,k)   COMMON /blck17/   in/out
 *   sap, smp [ncell][4], sion1o [ncell][4][9]   locals of stem_kpp carried between the calls
 *   nwarn  [ncell] or NULL    number of "aerosol growth" messages (x0 <= 0, str.f90:6027)
 * Numerics: every statement is evaluated in binary64 with the reference's operation
 * order (no FMA contraction), with two documented reassociations: (1) the particle
 * number of a bin, sap, is summed per dry class first and then over classes - the
 * reference keeps one running sum over all (ia, jt) (str.f90:5949), a 1600-long dependent
 * chain; (2) the transferred volumes vc are summed per water bin first.  Hence sap agrees
 * with the reference loop order to 1e-13, ff to 1e-11 and sl1/sion1 to 1e-11 relative
 * (measured: a few 1e-16).  libmistra_kpp_strict.so (-DKPP_STRICT) keeps the running sum
 * for sap and reproduces ff, sap, smp to the last bit; the tests check both.
 *
 * All functions return 0 or a negative MISTRA_KPP_E* code (mistra_kpp.h); the text is
 * available from mistra_kpp_last_error().  No CPU fallback.
 */
#ifndef MISTRA_BINS_H
#define MISTRA_BINS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MISTRA_NKC 4   /* global_params.f90: nkc */
#define MISTRA_J2 121  /* global_params.f90: j2  */
#define MISTRA_J6 55   /* global_params.f90: j6  */
#define MISTRA_LSP 9   /* str.f90:5837 lsp       */

/* The 2-D particle grid (SUBROUTINE grid, str.f90:1653-1705, 1882-1905; COMMON /cb50/,
 * /blck06/) and the switches stem_kpp reads.  kw and ka keep the reference's 1-based
 * meaning: water bins jt <= kw(ia) of dry class ia are "aerosol", the rest "droplet";
 * dry classes ia <= ka belong to chem bins 1/3, ia > ka to bins 2/4. */
typedef struct mistra_bins_grid {
  int32_t nka, nkt;      /* 70, 70 (<= 128 each)                                   */
  int32_t ka;            /* COMMON /blck06/ ka                                      */
  int32_t nkc_l;         /* config: number of chem bins in use (1..4)               */
  int32_t ial_first;     /* 1, or 2 when nuc .and. ifeed == 2 (str.f90:5925-5929)   */
  int32_t reserved;
  const int32_t *kw;     /* [nka]       COMMON /blck06/ kw                          */
  const double *en;      /* [nka]       dry aerosol mass per class [mg]             */
  const double *rq;      /* [nka][nkt]  = rq(nkt,nka) total particle radius [um]    */
} mistra_bins_grid;

/* HOST buffers (staged to the current CUDA device and back, synchronous). */
int mistra_bins_snapshot(const mistra_bins_grid *g, int64_t ncell, const double *ff,
                         const double *cm, const double *sion1, double *sap, double *smp,
                         double *sion1o, void *stream);
int mistra_bins_redistribute(const mistra_bins_grid *g, int64_t ncell, double *ff,
                             const double *cm, const double *cw, const double *sap,
                             const double *smp, const double *sion1o, double *sion1,
                             double *sl1, int32_t *nwarn, void *stream);

/* DEVICE buffers on the current device, asynchronous on `stream` (NULL = legacy default
 * stream).  The grid struct itself is read on the host; its kw/en/rq arrays are HOST
 * pointers (copied to the device once per distinct grid). */
int mistra_bins_snapshot_device(const mistra_bins_grid *g, int64_t ncell, const double *d_ff,
                                const double *d_cm, const double *d_sion1, double *d_sap,
                                double *d_smp, double *d_sion1o, void *stream);
int mistra_bins_redistribute_device(const mistra_bins_grid *g, int64_t ncell, double *d_ff,
                                    const double *d_cm, const double *d_cw, const double *d_sap,
                                    const double *d_smp, const double *d_sion1o, double *d_sion1,
                                    double *d_sl1, int32_t *d_nwarn, void *stream);

/* Kernels launched by the bins entries since load. */
int64_t mistra_bins_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
