/* mistra_drive.h - C ABI of the gather / scatter halves of gas_drive / aer_drive / tot_drive on the device
 * (part of libmistra_kpp.so).
 *
 * Rows a14 / a15 of the hot-path scope and the second half of "next" row N3: the copies
 * between the model arrays s1(j1,n), s3(j5,n), sl1(j2,nkc,n), sion1(j6,nkc,n) and the KPP vectors C =
 * (VAR, FIX) that /root/reference/src/gas.f:60-217, aer.f:59-246 and tot.f:59-982 make around
 * Update_RCONST_x / INTEGRATE_x for one layer k, here for a batch of layers whose arrays stay in device
 * memory between mistra_difc, mistra_konc, mistra_rconst_update_device, mistra_kpp_integrate_device and
 * mistra_bins_redistribute.  Device pointers only (a host caller has nothing to gain from it).
 *
 * gather (before the integration), per batch cell i with layer k = layer[i]:
 *   sl1(:,:,k) = max(0, sl1(:,:,k)); sion1(:,:,k) = max(0, sion1(:,:,k))   when clamp_liquid (aer.f:158-159)
 *   C(map_kpp(m)) = <map_arr(m)>(map_off(m), k)   for every map entry: s1 / s3 through gas_m2k_x / rad_m2k_x
 *                   (aer.f:146-151), sl1 / sion1 through aer_mk.dat and the l3 / l4 block of tot.f:256-597
 *   FIX(indf_O2) = 0.21*air, FIX(indf_H2O) = h2o, FIX(indf_N2) = 0.79*air, FIX(indf_H2Olc) = 55.55/cvvc or 0
 *                   (aer.f:153-170; the three constants are default-REAL literals there: f32_literals)
 *   entries of VAR no statement assigns keep what the caller left in var (the reference's C is a COMMON
 *   block that keeps the previous layer's value, SURVEY trap 8).
 * scatter (after it): <map_arr(m)>(map_off(m), k) = C(map_kpp(m)) (aer.f:233-245, aer_km.dat), then, with
 *   clip_negative, every element of the layer's four rows is clipped at 0 (kpp_driver, kpp.f90:4472-4477).
 * Pure copies: bit-exact.  Returns 0 or MISTRA_KPP_E* (mistra_kpp.h).  No CPU fallback. */
#ifndef MISTRA_DRIVE_H
#define MISTRA_DRIVE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mistra_drive_args {
  int32_t nvar, nfix;                 /* row lengths of var and fix (NVAR, NFIX of the mechanism) */
  int32_t j1, j5, j2, j6, nkc;        /* s1[nlayer][j1], s3[nlayer][j5], sl1[nlayer][nkc][j2], sion1[nlayer][nkc][j6] */
  int32_t nmap;
  const int32_t *map_kpp;             /* [nmap] 1-based index into C: <= nvar -> VAR, else FIX(idx - nvar) */
  const int32_t *map_arr;             /* [nmap] 0: s1, 1: s3, 2: sl1, 3: sion1 */
  const int32_t *map_off;             /* [nmap] 0-based offset inside the layer's row of that array */
  const int64_t *layer;               /* [ncell] row of the model arrays each batch cell stands for */
  double *s1, *s3, *sl1, *sion1;
  double *var, *fix;                  /* [ncell][nvar], [ncell][nfix] */
  /* gather only */
  int32_t clamp_liquid;
  int32_t f32_literals;
  int32_t indf_o2, indf_h2o, indf_n2; /* 1-based positions in FIX, 0 = not set */
  int32_t indf_h2ol[4];
  const double *air, *h2o;            /* [ncell] */
  const double *cvv;                  /* [ncell][4] cvv1..4, read where indf_h2ol != 0 */
  /* scatter only */
  int32_t clip_negative;
} mistra_drive_args;

int mistra_drive_gather_device(int64_t ncell, const mistra_drive_args *d_a, void *stream);
int mistra_drive_scatter_device(int64_t ncell, const mistra_drive_args *d_a, void *stream);

int64_t mistra_drive_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
