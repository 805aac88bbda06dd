// Mechanism 'a' device code: generated straight-line Fun/Jac/LU/solve + the Ros3 kernel.
#include "kpp_batch.h"
#include "_gen/mech_a.cuh"
#define MECH_NS mech_a
#define ROS3_KERNEL ros3_kernel_a
#define ROS3_LAUNCH ros3_launch_a
#include "ros3_kernel.inc"

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
                                   kpp_onchip_info_a()};
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
