// Mechanism 'g' device code: generated straight-line Fun/Jac/LU/solve + the Ros3 kernel.
#include "kpp_batch.h"
#include "_gen/mech_g.cuh"
#define MECH_NS mech_g
#define ROS3_KERNEL ros3_kernel_g
#define ROS3_LAUNCH ros3_launch_g
#include "ros3_kernel.inc"

// On-chip kernel: one thread block per cell (mechgen/onchip.py, ros3_onchip.inc)
#include "_gen/onchip_g.cuh"
namespace oc_g {
__constant__ double c_lit[NLIT];
}
#define OC_NS oc_g
#define OC_KERNEL ros3_onchip_g
#define OC_CTAS 16
#define OC_W 1
#include "ros3_onchip.inc"
extern "C" const unsigned short mistra_oc_tables_g[];
extern "C" const size_t mistra_oc_tables_g_count;
namespace oc_g {
static cudaError_t launch(const KppBatch &b, int blocks, cudaStream_t st)
{
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ros3_onchip_g, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DOUBLES * 8);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ros3_onchip_g, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  ros3_onchip_g<<<blocks, NT, SMEM_DOUBLES * 8, st>>>(b);
  return cudaGetLastError();
}
static cudaError_t set_lit(const double *h, cudaStream_t st)
{
  return cudaMemcpyToSymbolAsync(c_lit, h, sizeof(double) * NLIT, 0, cudaMemcpyHostToDevice, st);
}
}  // namespace oc_g

namespace mech_g {
static cudaError_t set_coef(const double *h, cudaStream_t st)
{
  return cudaMemcpyToSymbolAsync(c_coef, h, sizeof(double) * NCOEF, 0, cudaMemcpyHostToDevice, st);
}
}  // namespace mech_g

const KppMechInfo *kpp_mech_info_g()
{
  using namespace mech_g;
  static const KppMechInfo info = {NVAR, NFIX, NREACT, LU_NONZERO, NSLOT, NCOEF, coef_literals,
                                   (const void *)ros3_kernel_g, ros3_launch_g, set_coef,
                                   (const void *)oc_g::ros3_onchip_g, oc_g::launch, oc_g::set_lit, mistra_oc_tables_g,
                                   mistra_oc_tables_g_count, oc_g::coef_literals, oc_g::NLIT, oc_g::NT,
                                   oc_g::SMEM_DOUBLES * 8, oc_g::T, OC_CTAS};
  return &info;
}

#ifdef KPP_PHASE_TIMERS
extern "C" int mistra_kpp_phase_g(unsigned long long *out, int reset)
{
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, mech_g::g_phase, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(mech_g::g_phase, z, sizeof z); }
  return (int)e;
}
#endif
