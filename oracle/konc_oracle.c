/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h for who may load this).
 *
 * Restatement of SUBROUTINE konc (/root/reference/src/kpp.f90:3370-3585): after kon has changed
 * the particle spectrum, dissolved species follow the liquid volume that moved between the
 * aerosol bins (1, 2) and the droplet bins (3, 4).  Statement order of the reference; the
 * warnings it prints (kpp.f90:3443-3445, 3486-3489, 3511-3513, 3541-3544) are counted.
 * Parity unpinned by the reference (it ships no tests or fixtures for this routine).
 *
 * Arrays are the reference's COMMON arrays with the layer index last:
 *   vol1_a, vol1_d, part_o_a, part_o_d, part_n_a, part_n_d [ncell][nka]   /blck07/, /blck08/
 *   vol2, pntot [ncell][4]
 *   sl1 [ncell][4][j2], sion1 [ncell][4][j6]                              /blck17/
 *   warn [ncell][3]: number of classes with |dp_a + dp_d| > 1e-10, delta < 0, delta > 1
 */
#include <math.h>
#include <stdint.h>

static void move_frac(double *s, int n, int stride_bin, int ii, int jj, double delta)
{
  /* kpp.f90:3491-3500 */
  for (int l = 0; l < n; ++l) {
    double *a = s + (ii - 1) * stride_bin + l, *b = s + (jj - 1) * stride_bin + l;
    const double del = *a * delta;
    *a = fmax(0.0, *a - del);
    *b = fmax(0.0, *b + del);
  }
}

static void konc_layer(int nka, int ka, int j2, int j6, const double *vol1_a, const double *vol1_d,
                       const double *part_o_a, const double *part_o_d, const double *part_n_a,
                       const double *part_n_d, const double *vol2, const double *pntot, double *sl1,
                       double *sion1, int32_t *warn)
{
  warn[0] = warn[1] = warn[2] = 0;
  for (int ia = 0; ia < nka; ++ia) {
    /* small classes (ia <= ka) exchange between bins 1 and 3, large ones between 2 and 4
       (kpp.f90:3438-3503 and 3506-3562 are the same statements with the bin numbers changed) */
    const int ba = ia < ka ? 1 : 2, bd = ba + 2;
    const double dp_a = part_o_a[ia] - part_n_a[ia];
    const double dp_d = part_o_d[ia] - part_n_d[ia];
    if (fabs(dp_a + dp_d) > 1.e-10) warn[0]++;
    const int ii = (dp_a >= 1.e-10) ? ba : bd;
    const double xs = (fabs(dp_a) < 1.e-10) ? 0.0 : 1.0;
    int jj;
    double delta;
    if (ii == ba) {
      jj = bd;
      if (vol2[ii - 1] > 0. && part_o_a[ia] > 0.) delta = vol1_a[ia] / vol2[ii - 1] * dp_a / part_o_a[ia] * xs;
      else delta = 0.0;
    } else {
      jj = ba;
      if (vol2[ii - 1] > 0. && part_o_d[ia] > 0.) delta = vol1_d[ia] / vol2[ii - 1] * dp_d / part_o_d[ia] * xs;
      else delta = 0.0;
    }
    if (delta < 0.0) warn[1]++;
    else if (delta > 1.0) warn[2]++;
    else if (delta > 0.0) {
      move_frac(sl1, j2, j2, ii, jj, delta);
      move_frac(sion1, j6, j6, ii, jj, delta);
    }
  }
  /* too few droplets left: everything back to the aerosol bin (kpp.f90:3566-3586) */
  for (int l = 0; l < j2; ++l) {
    if (pntot[2] < 1.e-7) { sl1[l] = sl1[l] + fmax(0.0, sl1[2 * j2 + l]); sl1[2 * j2 + l] = 0.0; }
    if (pntot[3] < 1.e-7) { sl1[j2 + l] = sl1[j2 + l] + fmax(0.0, sl1[3 * j2 + l]); sl1[3 * j2 + l] = 0.0; }
  }
  for (int l = 0; l < j6; ++l) {
    if (pntot[2] < 1.e-7) { sion1[l] = sion1[l] + fmax(0.0, sion1[2 * j6 + l]); sion1[2 * j6 + l] = 0.0; }
    if (pntot[3] < 1.e-7) { sion1[j6 + l] = sion1[j6 + l] + fmax(0.0, sion1[3 * j6 + l]); sion1[3 * j6 + l] = 0.0; }
  }
}

void konc_oracle(int64_t ncell, int nka, int ka, int j2, int j6, const double *vol1_a, const double *vol1_d,
                 const double *part_o_a, const double *part_o_d, const double *part_n_a,
                 const double *part_n_d, const double *vol2, const double *pntot, double *sl1, double *sion1,
                 int32_t *warn)
{
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncell; ++c)
    konc_layer(nka, ka, j2, j6, vol1_a + c * nka, vol1_d + c * nka, part_o_a + c * nka, part_o_d + c * nka,
               part_n_a + c * nka, part_n_d + c * nka, vol2 + c * 4, pntot + c * 4, sl1 + c * 4 * j2,
               sion1 + c * 4 * j6, warn + c * 3);
}
