/* mistra_driver.h - C ABI of the layer loop of SUBROUTINE kpp_driver on the device (part of libmistra_kpp.so).
 *
 * Row a15 of the hot-path scope: /root/reference/src/kpp.f90:4305-4470 - what kpp_driver does for
 * every layer k = n_min .. n_max before it calls gas_drive / aer_drive / tot_drive: the clip of s1 / s3 (4305-4306),
 * the layer's scalars of COMMON /cb_1/ (te, air_cc, h2oppm, pk; 4315-4321), air and h2o in mol/m3, the conversion
 * factors cvv1..4 (4327-4330), the photolysis rates ph_rat averaged from the two bounding levels or zero at night
 * (4344-4362), the switches xhal, xiod (4365-4371), xliq1..4 from the water volume of the four bins (4374-4390), the
 * bookkeeping of COMMON /kpp_l1/ cloud (4392-4412), xhet1, xhet2 (4435-4438), the Eulerian advection source on s1
 * (4441-4449), and the choice of the mechanism (4452-4468).  Instead of calling a mechanism per layer, the layers are
 * sorted into three ascending lists (gas, aer, tot), which mistra_drive_gather / mistra_rconst_update_device /
 * mistra_kpp_integrate_device / mistra_drive_scatter then take as batches: with this entry a chemistry step needs no
 * host-built per-layer input.  chamber mode (4299-4304, 4331-4334, 4360-4362, 4414-4433) is not covered and stays
 * host code; what precedes the loop (solar angle u0, 4280-4287; mass_ch) is a handful of scalars per call and is an
 * input here.
 *
 * Arrays (level index last in Fortran = slowest here, column slowest of all; level 1 of the reference = index 0;
 * a layer's row in the outputs is L = col * n + (k - 1)):
 *   u0 [ncol]                           as kpp.f90:4280-4287 leaves it
 *   t, p, rho [ncol][n]  COMMON /cb53/  cm3, am3 [ncol][n]  /blck01/   xm1 [ncol][n]  /cb54/
 *   conv2 [ncol][n][nkc] /blck13/       cm [ncol][n][nkc] /blck12/     cloud [ncol][n][nkc] int32 IN/OUT /kpp_l1/
 *   photol_j [ncol][n][nphrxn]          /band_rat/
 *   s1 [ncol][n][j1], s3 [ncol][n][j5]  gas_common, IN/OUT (clip; advection source on s1); either may be NULL
 *   adv_row [nadv]: 0-based row of s1 of every advected species = ind_gas_rev(nindadv(j)) - 1, or -1 where
 *   nindadv(j) = 0;  xadv [nadv]
 * Outputs, rows of the levels n_min .. n_max only (the others are left as they are, mech = -1):
 *   cb1 [ncol*n][4] = air_cc, te, h2oppm, pk      scal [ncol*n][13] = conv1, xhal, xiod, xhet1, xhet2, xliq1..4,
 *   cvv1..4 (the rows mistra_rate_inputs takes, mistra_rconst.h)      ph_rat [ncol*n][nphrxn]
 *   air, h2o [ncol*n]   cvv [ncol*n][4]   mech [ncol*n] int32 (0 gas, 1 aer, 2 tot, -1 not integrated)
 *   layers [3][ncol*n] int64: the rows L of every mechanism in ascending order; count [3] int64
 * Numerics: binary64, the reference's expressions and operation order: bit-identical.
 * Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).  No CPU fallback. */
#ifndef MISTRA_DRIVER_H
#define MISTRA_DRIVER_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mistra_driver_args {
  int32_t n, nf, nkc, nphrxn;            /* global_params: n, nf, nkc (4), nphrxn (47) */
  int32_t halo, iod, lpBuys13_0D, neula; /* config.f90 switches (neula = 0: Eulerian advection source on) */
  int32_t box, n_bl;                     /* box run: only level n_bl (1-based) is integrated */
  int32_t kinv, nadv, j1, j5;
  double dt_ch;
  const double *u0;
  const double *t, *p, *rho, *cm3, *am3, *xm1;
  const double *conv2, *cm;
  int32_t *cloud;
  const double *photol_j;
  const int32_t *adv_row;
  const double *xadv;
  double *s1, *s3;
  double *cb1, *scal, *ph_rat, *air, *h2o, *cvv;
  int32_t *mech;
  int64_t *layers, *count;
} mistra_driver_args;

/* Every pointer is a DEVICE pointer on the current device; asynchronous on `stream`. */
int mistra_driver_layers_device(int64_t ncol, const mistra_driver_args *d_a, void *stream);
/* HOST buffers (staged to the current device and back; synchronous). */
int mistra_driver_layers(int64_t ncol, const mistra_driver_args *a, void *stream);

int64_t mistra_driver_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
