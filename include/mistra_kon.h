/* mistra_kon.h - C ABI of the condensation / evaporation step on the 2-D particle grid
 * (part of libmistra_kpp.so).
 *
 * Replaces, for all humid layers of one or many columns at once, the call
 *   call subkon (dt, ffk, totr, dfdt, feualt, pp, to, tn, xm1o, xm1n, kr)
 * inside the layer loop of SUBROUTINE kon (/root/reference/src/str.f90:4615-4721, call at
 * 4705), i.e. SUBROUTINE subkon (str.f90:4987-5204) with the advection scheme
 * SUBROUTINE advec (str.f90:5321-5516) and the functions diff_wat_vap, therm_conduct_air,
 * xl21, p21 (str.f90:5210-5315, 7640-7693): growth-rate coefficients cd, cr, sr over the
 * (water mass x dry mass) grid with the 18-band radiative term, then up to 10 secant
 * iterations on the mean saturation ratio, each advecting the 70 rows of the particle
 * spectrum along the water-mass axis with Bott's positive-definite scheme.
 *
 * The caller keeps the rest of kon: the dry branch (equil for feu < 0.7), the bin sums for
 * konc, and the write-back of t, xm1, feu, xm2 (str.f90:4708-4721) from the returned
 * to, xm1o and ffk.
 *
 * Arrays (C row-major = Fortran column-major, layer index last):
 *   ffk    [ncell][nka][nkt] = ff(nkt,nka,k)      particles cm^-3, in/out
 *   totr   [ncell][18]       = totrad(1:mb,k)     radiative flux per band
 *   dfdt, feualt, pp, tn, xm1n [ncell]            dfddt(k), feu(k), p(k), t(k), xm1(k)
 *   to, xm1o [ncell]                              talt(k), xm1a(k) in; new t, xm1 out
 *   kr     [ncell]                                nar(k): aerosol type 1..3 (qabs table)
 *   status [ncell] or NULL   iterations used (1..10); -1 = "no convergence of condensation
 *                            iteration" (the reference warns and carries on with the last
 *                            iterate, so do we); -2 = advec left the grid (the reference
 *                            aborts: "SR advec: error with k_high or k_low")
 * Numerics: binary64, statement order of the reference; exp/pow come from the CUDA math
 * library (<= 1-2 ulp from the host libm), and the liquid-water change dwsum is summed per
 * dry class first - results agree with the reference loop order to ~1e-12 relative, not to
 * the bit.  Returns 0 or a negative MISTRA_KPP_E* code (mistra_kpp.h).  No CPU fallback.
 */
#ifndef MISTRA_KON_H
#define MISTRA_KON_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MISTRA_MB 18        /* global_params.f90: mb        */
#define MISTRA_JPTAERRAD 3  /* global_params.f90: jptaerrad */

/* COMMON /cb44/, /cb49/, /cb50/, /cb51/ as subkon reads them (HOST pointers; uploaded once
 * per distinct grid). */
typedef struct mistra_kon_grid {
  int32_t nka, nkt;      /* <= 128 each                                              */
  double a0m;            /* COMMON /cb44/ a0m = 152200/(r1*rhow) (str.f90:1317)       */
  double dlne;           /* COMMON /cb51/ dlne = ln(10)*dlgew (str.f90:1683)          */
  const double *en;      /* [nka]  dry aerosol mass [mg]                              */
  const double *rn;      /* [nka]  dry radius [um]                                    */
  const double *b0m;     /* [nka]  COMMON /cb44/ solute term (str.f90:1398)           */
  const double *ew;      /* [nkt]  upper water-mass limit of bin jt [mg]              */
  const double *e;       /* [nkt]  mean water mass of bin jt [mg]                     */
  const double *dew;     /* [nkt]  bin width [mg]                                     */
  const double *rw;      /* [nka][nkt] = rw(nkt,nka) radius at the upper bin limit [um] */
  const double *qabs;    /* [3][nka][nkt][18] = qabs(18,nkt,nka,jptaerrad)            */
  /* only read by mistra_kon_layers (may be NULL / 0 for mistra_kon_subkon): */
  const int32_t *kw;     /* [nka]  COMMON /blck06/ kw (1-based limit aerosol|droplet)   */
  const double *rq;      /* [nka][nkt] = rq(nkt,nka) radius at the bin mean [um]        */
  int32_t ka;            /* COMMON /blck06/ ka                                          */
  int32_t reserved;
} mistra_kon_grid;

/* HOST buffers (staged to the current device and back; synchronous). */
int mistra_kon_subkon(const mistra_kon_grid *g, int64_t ncell, double dt, double *ffk,
                      const double *totr, const double *dfdt, const double *feualt,
                      const double *pp, double *to, const double *tn, double *xm1o,
                      const double *xm1n, const int32_t *kr, int32_t *status, void *stream);

/* DEVICE buffers on the current device, asynchronous on `stream` (NULL = legacy default). */
int mistra_kon_subkon_device(const mistra_kon_grid *g, int64_t ncell, double dt, double *d_ffk,
                             const double *d_totr, const double *d_dfdt, const double *d_feualt,
                             const double *d_pp, double *d_to, const double *d_tn,
                             double *d_xm1o, const double *d_xm1n, const int32_t *d_kr,
                             int32_t *d_status, void *stream);

/* The whole layer loop of SUBROUTINE kon (str.f90:4615-4772) for ncell layers at once:
 * per layer, with chem, the bin sums before the step (vol1_a/d, part_o_a/d, vol2;
 * 4625-4659); then either the dry branch feu(k) < 0.7 - feu from xm1 and SUBROUTINE equil
 * case 1 (Koehler equilibrium size of every dry class by FUNCTION rgl, str.f90:4801-4981,
 * 2164-2251), which also updates xm2 - or the humid branch: subkon as above followed by the
 * write-back t = talt = to, xm1 = xm1a = xm1o, feu, dfddt, xm2, dtcon (4708-4721); then, with
 * chem, the sums after the step (part_n_a/d, pntot; 4724-4770).  What stays on the host:
 * the cloud base/top search (4775-4782, a scan over k) and konc.
 * Arrays are the reference's COMMON arrays with the layer index last; all [ncell] unless
 * noted.  With chem == 0 the eight sum arrays may be NULL.
 * status: 0 = dry branch taken, 1..10 / -1 / -2 as for mistra_kon_subkon. */
typedef struct mistra_kon_state {
  double *ff;                      /* [ncell][nka][nkt] COMMON /cb52/  in/out            */
  double *t, *talt;                /* COMMON /cb53/   in/out                              */
  double *xm1, *xm1a, *feu, *dfddt;/* COMMON /cb54/   in/out                              */
  double *xm2;                     /* COMMON /cb54/   out                                 */
  double *dtcon;                   /* COMMON /cb48/   out                                 */
  const double *p;                 /* COMMON /cb53/                                       */
  const double *totrad;            /* [ncell][18] COMMON /cb11/                           */
  const int32_t *nar;              /* COMMON /cb52/                                       */
  double *vol1_a, *vol1_d, *part_o_a, *part_o_d, *part_n_a, *part_n_d; /* [ncell][nka]   */
  double *vol2, *pntot;            /* [ncell][4]  COMMON /blck07/, /blck08/               */
  int32_t *status;                 /* or NULL                                             */
} mistra_kon_state;

/* HOST buffers (synchronous) / DEVICE buffers (asynchronous on `stream`). */
int mistra_kon_layers(const mistra_kon_grid *g, int64_t ncell, double dt, int chem,
                      const mistra_kon_state *s, void *stream);
int mistra_kon_layers_device(const mistra_kon_grid *g, int64_t ncell, double dt, int chem,
                             const mistra_kon_state *d_s, void *stream);

int64_t mistra_kon_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
