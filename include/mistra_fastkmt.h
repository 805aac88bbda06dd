/* mistra_fastkmt.h - C ABI of SUBROUTINE fast_k_mt_a / fast_k_mt_t on the device (part of
 * libmistra_kpp.so).
 *
 * Second piece of "next" row N2 of the hot-path scope (the liq_parm chain): replaces
 * the layer loop of SUBROUTINE fast_k_mt_a (/root/reference/src/kpp.f90:2683-2947) and of
 * fast_k_mt_t (kpp.f90:2421-2676; the same body with the tot species indices and kc = 1..nkc
 * instead of 1..nkc_l) for all layers at once: the gas <-> aqueous mass-transfer coefficients
 * xkmt(lex(l),kc,k) of the nx = 50 exchanged species (Schwartz 1986, integrated over the 2-D particle
 * spectrum ff instead of a mean radius) and the liquid-water-weighted sedimentation velocity
 * vt(kc,k) of each chemistry bin (FUNCTION vterm, str.f90:2793-2864).  Called from liq_parm every
 * 120 s or when a bin is switched on (kpp.f90:611-621).  Reads the same ff that mistra_cwrc,
 * mistra_kon_layers and mistra_bins_redistribute keep on the device, and the cw / cm that
 * mistra_cwrc writes.
 *
 * Arrays (layer index last in Fortran = first here):
 *   ff    [ncell][nka][nkt]   COMMON /cb52/           freep, t, p [ncell]  (liq_parm: freep = 2.28e-5 t/p)
 *   cw, cm [ncell][nkc]       COMMON /blck12/         alpha, vmean [ncell][nspec]  COMMON /kpp_2aer/ (/kpp_2tot/)
 *   xkmt  [ncell][nkc][nspec] COMMON /kpp_laer/ (/kpp_ltot/)   IN/OUT: only the entries the reference
 *                              assigns are written (species of lex, bins with cm > 0 and cw > 0)
 *   vt    [ncell][nkc]        COMMON /kpp_vt/         IN/OUT: written where cw > 0
 *   lex   [nx] 1-based KPP species indices (the DATA statement kpp.f90:2794-2803 / 2532-2541)
 *   kw [nka] (1-based limit), rq [nka][nkt] in um     COMMON /blck06/, /cb50/
 * Numerics: binary64, every term with the reference's expression and operation order, no FMA
 * contraction; log / exp of vterm are CUDA's (<= 1 ulp, as any libm).  Each of the (nx + 1) * nkc
 * sums is formed as per-warp partial sums over an interleaved point order, added in warp order (the
 * reference keeps one running sum over (ia, jt), a chain of up to 2660 additions): all terms are
 * non-negative, results agree with the reference order to ~1e-14 relative (tests: 1e-13).
 * Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).  No CPU fallback. */
#ifndef MISTRA_FASTKMT_H
#define MISTRA_FASTKMT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MISTRA_FASTKMT_MAXNX 64

typedef struct mistra_fastkmt_args {
  int32_t nka, nkt, ka;
  int32_t ial;          /* 2 when ifeed == 2, else 1 (kpp.f90:2870-2876) */
  int32_t nkc;          /* second dimension of cw, cm, xkmt, vt (global_params nkc = 4) */
  int32_t nkc_l;        /* bins looped over: config nkc_l for fast_k_mt_a, nkc for fast_k_mt_t */
  int32_t nspec;        /* NSPEC of the mechanism (aer_Parameters.h / tot_Parameters.h) */
  int32_t nx;           /* number of exchanged species, <= MISTRA_FASTKMT_MAXNX */
  const int32_t *lex;
  const int32_t *kw;
  const double *rq;
  const double *ff, *freep, *t, *p, *cw, *cm, *alpha, *vmean;
  double *xkmt, *vt;
} mistra_fastkmt_args;

/* HOST buffers (staged to the current device and back; synchronous). */
int mistra_fastkmt(int64_t ncell, const mistra_fastkmt_args *a, void *stream);
/* Every pointer is a DEVICE pointer on the current device; asynchronous on `stream`. */
int mistra_fastkmt_device(int64_t ncell, const mistra_fastkmt_args *d_a, void *stream);

int64_t mistra_fastkmt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
