// Mechanism 'a' device code: generated straight-line Fun/Jac/LU/solve + the Ros3 kernel.
#include "kpp_batch.h"
#include "_gen/mech_a.cuh"
#define MECH_NS mech_a
#define ROS3_KERNEL ros3_kernel_a
#define ROS3_LAUNCH ros3_launch_a
#include "ros3_kernel.inc"

// On-chip kernel: one thread block per cell (mechgen/onchip.py, ros3_onchip.inc)
#include "_gen/onchip_a.cuh"
namespace oc_a {
__constant__ double c_lit[NLIT];
}
#define OC_NS oc_a
#define OC_KERNEL ros3_onchip_a
#define OC_CTAS 5
#define OC_W 2
#include "ros3_onchip.inc"
extern "C" const unsigned short mistra_oc_tables_a[];
extern "C" const size_t mistra_oc_tables_a_count;
namespace oc_a {
static cudaError_t launch(const KppBatch &b, int blocks, cudaStream_t st)
{
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ros3_onchip_a, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DOUBLES * 8);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ros3_onchip_a, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  ros3_onchip_a<<<blocks, NT, SMEM_DOUBLES * 8, st>>>(b);
  return cudaGetLastError();
}
static cudaError_t set_lit(const double *h, cudaStream_t st)
{
  return cudaMemcpyToSymbolAsync(c_lit, h, sizeof(double) * NLIT, 0, cudaMemcpyHostToDevice, st);
}
}  // namespace oc_a

namespace mech_a {
static cudaError_t set_coef(const double *h, cudaStream_t st)
{
  return cudaMemcpyToSymbolAsync(c_coef, h, sizeof(double) * NCOEF, 0, cudaMemcpyHostToDevice, st);
}
}  // namespace mech_a

const KppMechInfo *kpp_mech_info_a()
{
  using namespace mech_a;
  static const KppMechInfo info = {NVAR, NFIX, NREACT, LU_NONZERO, NSLOT, NCOEF, coef_literals,
                                   (const void *)ros3_kernel_a, ros3_launch_a, set_coef,
                                   (const void *)oc_a::ros3_onchip_a, oc_a::launch, oc_a::set_lit, mistra_oc_tables_a,
                                   mistra_oc_tables_a_count, oc_a::coef_literals, oc_a::NLIT, oc_a::NT,
                                   oc_a::SMEM_DOUBLES * 8, oc_a::T, OC_CTAS};
  return &info;
}

#ifdef KPP_PHASE_TIMERS
extern "C" int mistra_kpp_phase_a(unsigned long long *out, int reset)
{
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, mech_a::g_phase, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(mech_a::g_phase, z, sizeof z); }
  return (int)e;
}
#endif
