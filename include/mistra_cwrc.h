/* mistra_cwrc.h - C ABI of SUBROUTINE cw_rc on the device (part of libmistra_kpp.so).
 *
 * First piece of "next" row N2 of the hot-path scope (the liq_parm chain):
 * replaces the layer loop of SUBROUTINE cw_rc (/root/reference/src/kpp.f90:2152-2414, called
 * from liq_parm, kpp.f90:565) for all layers at once - the liquid water content cw, mean radius
 * rc and water volume cm of the four chemistry bins as sums over the 2-D particle spectrum
 * ff(jt,ia,k), and from them the chemistry switches cm / conv2 (kpp.f90:2366-2410) that
 * kpp_driver turns into xliq1..4 and cvv1..4 (kpp.f90:4327-4438).  Reads the same ff that
 * mistra_kon_layers and mistra_bins_redistribute keep on the device.
 *
 * Arrays (layer index last in Fortran = first here):
 *   ff    [ncell][nka][nkt]  COMMON /cb52/          feu   [ncell]     COMMON /cb54/
 *   cloud [ncell][4] int32   COMMON /kpp_l1/ (0 = .false.)
 *   rc, cw, cm, conv2 [ncell][4]   COMMON /blck11/, /blck12/, /blck13/   out
 *   kw [nka] (1-based limit), e [nkt], rq [nka][nkt]   COMMON /blck06/, /cb50/
 * Numerics: binary64, the reference's expressions without FMA contraction; the twelve sums are
 * formed as 256 interleaved partial sums (dry classes ia = w mod 8, water bins jt = l mod 32 per
 * warp w and lane l) combined by a fixed butterfly over the lanes and then over the warps in order
 * (the reference keeps one running sum over all (ia, jt) of a bin, a chain of up to 2660 additions): all terms are non-negative, results agree
 * with the reference order to ~1e-14 relative (tests: 1e-13).
 * Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).  No CPU fallback. */
#ifndef MISTRA_CWRC_H
#define MISTRA_CWRC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mistra_cwrc_args {
  int32_t nka, nkt, ka;
  int32_t ial;                                      /* 2 when ifeed == 2, else 1 (kpp.f90:2253-2257) */
  double xcryssulf, xcrysss, xdelisulf, xdeliss;    /* COMMON /kpp_crys/ (set in initc, kpp.f90:319-324) */
  const int32_t *kw;
  const double *e, *rq;
  const double *ff, *feu;
  const int32_t *cloud;
  double *rc, *cw, *cm, *conv2;
} mistra_cwrc_args;

/* HOST buffers (staged to the current device and back; synchronous). */
int mistra_cwrc(int64_t ncell, const mistra_cwrc_args *a, void *stream);
/* Every pointer is a DEVICE pointer on the current device; asynchronous on `stream`. */
int mistra_cwrc_device(int64_t ncell, const mistra_cwrc_args *d_a, void *stream);

int64_t mistra_cwrc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
