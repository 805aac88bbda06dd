// Mechanism 'g' device code: generated straight-line Fun/Jac/LU/solve + the Ros3 kernel.
#include "kpp_batch.h"
#include "_gen/mech_g.cuh"
#define MECH_NS mech_g
#define ROS3_KERNEL ros3_kernel_g
#define ROS3_LAUNCH ros3_launch_g
#include "ros3_kernel.inc"

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
                                   kpp_onchip_info_g()};
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
