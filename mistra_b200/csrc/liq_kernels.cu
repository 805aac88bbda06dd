// henry_x, v_mean_x, st_coeff_x, equil_co_x on the device (include/mistra_liq.h): one thread per layer runs the
// generated statements of the four routines (csrc/liq_tables.h, ~300 assignments, ~150 exp / sqrt) into rows staged in
// LOCAL arrays would spill - instead every thread writes straight to its rows of the outputs (zero-filled first).
#include "liq_tables.h"
#include "../../include/mistra_kpp.h"
#include "../../include/mistra_liq.h"

#include <cuda_runtime.h>

#include <atomic>
#include <string>

int mistra_internal_fail(int code, const std::string &msg);

namespace {
std::atomic<long long> g_launches{0};

template <int MECH>
__global__ void __launch_bounds__(128) liq_tables_kernel(long long ncell, mistra_liq_args a)
{
  constexpr int NKC = MECH == 1 ? 2 : 4, NSPEC = MECH == 1 ? 262 : 424;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  LiqLayer L;
  L.tk = a.t[c];
  L.conv2 = a.conv2 + c * NKC;
  L.xgamma = a.xgamma + c * NKC * a.j6;
  L.lpjoyce14bc = a.lpjoyce14bc;
  L.lpbuxmann15alph = a.lpbuxmann15alph;
  L.j6 = a.j6;
  L.f32 = a.f32_literals;
  for (int kc = 0; kc < 4; ++kc) L.an2o5[kc] = 0.0;
  if (a.lpjoyce14bc)
    for (int kc = 0; kc < NKC; ++kc)
      L.an2o5[kc] = liq_a_n2o5(a.cw + c * NKC, a.cm + c * NKC, kc, a.sion1_13_14[(c * NKC + kc) * 2], a.sion1_13_14[(c * NKC + kc) * 2 + 1]);
  L.henry = a.henry + c * NSPEC;
  L.vmean = a.vmean + c * NSPEC;
  L.alpha = a.alpha + c * NSPEC;
  L.xkef = a.xkef + c * NKC * NSPEC;
  L.xkeb = a.xkeb + c * NKC * NSPEC;
  liq_tables_layer(MECH, L);
}
}  // namespace

extern "C" {

int mistra_liq_tables_device(int mech, int64_t ncell, const mistra_liq_args *a, void *stream)
{
  if ((mech != 1 && mech != 2) || !a || ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "liq tables: mech must be 1 (aer) or 2 (tot)");
  if (ncell == 0) return 0;
  if (!a->t || !a->conv2 || !a->xgamma || !a->henry || !a->vmean || !a->alpha || !a->xkef || !a->xkeb)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "liq tables: null array");
  if (a->lpjoyce14bc && (!a->cw || !a->cm || !a->sion1_13_14)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "liq tables: lpJoyce14bc needs cw, cm, sion1");
  if (a->j6 < 38) return mistra_internal_fail(MISTRA_KPP_EINVAL, "liq tables: j6 too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nkc = mech == 1 ? 2 : 4, nspec = mech == 1 ? 262 : 424;
  const size_t n = (size_t)ncell;
  cudaError_t e;
  if ((e = cudaMemsetAsync(a->henry, 0, n * nspec * 8, st)) != cudaSuccess || (e = cudaMemsetAsync(a->vmean, 0, n * nspec * 8, st)) != cudaSuccess ||
      (e = cudaMemsetAsync(a->alpha, 0, n * nspec * 8, st)) != cudaSuccess || (e = cudaMemsetAsync(a->xkef, 0, n * nkc * nspec * 8, st)) != cudaSuccess ||
      (e = cudaMemsetAsync(a->xkeb, 0, n * nkc * nspec * 8, st)) != cudaSuccess)
    return mistra_internal_fail(MISTRA_KPP_ECUDA, std::string("liq tables memset: ") + cudaGetErrorString(e));
  const unsigned blocks = (unsigned)((ncell + 127) / 128);
  if (mech == 1) liq_tables_kernel<1><<<blocks, 128, 0, st>>>(ncell, *a);
  else liq_tables_kernel<2><<<blocks, 128, 0, st>>>(ncell, *a);
  if ((e = cudaGetLastError()) != cudaSuccess) return mistra_internal_fail(MISTRA_KPP_ECUDA, std::string("liq_tables_kernel: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return 0;
}

int64_t mistra_liq_launch_count(void) { return g_launches.load(); }

}  // extern "C"

#include "liq_host.inc"
