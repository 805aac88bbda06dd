/* mistra_konc.h - C ABI of SUBROUTINE konc on the device (part of libmistra_kpp.so).
 *
 * "Next" row N3 of the hot-path scope: replaces, for all layers of one or many
 * columns at once, the layer loop of SUBROUTINE konc (/root/reference/src/kpp.f90:3370-3585,
 * called from kon, /root/reference/src/str.f90:4788): after the condensation step has moved
 * particles between the aerosol part (water bins below kw(ia)) and the droplet part of every
 * dry class, the dissolved species sl1 / sion1 of chemistry bins 1, 2 (aerosol) and 3, 4
 * (droplets) follow the liquid volume that went with them; if fewer than 1e-7 droplets cm^-3
 * remain in bin 3 (4) its whole content returns to bin 1 (2).
 *
 * Inputs are exactly what mistra_kon_layers(chem = 1) leaves behind (include/mistra_kon.h), so
 * kon -> konc -> stem_kpp can stay on the device.  Arrays (layer index last in Fortran =
 * first here):
 *   vol1_a, vol1_d, part_o_a, part_o_d, part_n_a, part_n_d  [ncell][nka]   COMMON /blck07/, /blck08/
 *   vol2, pntot  [ncell][4]
 *   sl1   [ncell][4][j2]   COMMON /blck17/   in/out   (j2 = 121 in global_params.f90)
 *   sion1 [ncell][4][j6]                     in/out   (j6 = 55)
 *   warn  [ncell][3] or NULL: per layer the number of dry classes for which the reference
 *         prints "Warning SR konc dp_1 > dp_3" (kpp.f90:3443), "... s <" / "l <" (delta < 0,
 *         3486, 3541) and "... s >" / "l >" (delta > 1, 3488, 3543); like the reference the
 *         class is then left unchanged.
 * ka = COMMON /blck06/ ka: classes 1..ka exchange between bins 1 and 3, ka+1..nka between 2 and 4.
 * Numerics: binary64, statement order of the reference, no FMA contraction: bit-identical to
 * the CPU restatement in oracle/konc_oracle.c.  Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).
 * No CPU fallback. */
#ifndef MISTRA_KONC_H
#define MISTRA_KONC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mistra_konc_args {
  int32_t nka, ka, j2, j6;
  const double *vol1_a, *vol1_d, *part_o_a, *part_o_d, *part_n_a, *part_n_d;
  const double *vol2, *pntot;
  double *sl1, *sion1;
  int32_t *warn;
} mistra_konc_args;

/* HOST buffers (staged to the current device and back; synchronous). */
int mistra_konc(int64_t ncell, const mistra_konc_args *a, void *stream);
/* DEVICE buffers on the current device, asynchronous on `stream` (NULL = legacy default). */
int mistra_konc_device(int64_t ncell, const mistra_konc_args *d_a, void *stream);

int64_t mistra_konc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
