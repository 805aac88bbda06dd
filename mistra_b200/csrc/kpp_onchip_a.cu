// Mechanism 'a': the on-chip Ros3 kernel (one thread block per cell), its own translation unit.
#include "kpp_onchip.h"
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


const KppOnchipInfo *kpp_onchip_info_a()
{
  using namespace oc_a;
  static const KppOnchipInfo info = {(const void *)ros3_onchip_a, launch, set_lit, mistra_oc_tables_a,
                                     mistra_oc_tables_a_count, coef_literals, NLIT, NT, SMEM_DOUBLES * 8, T, OC_CTAS};
  return &info;
}
