// Mechanism 'g': the on-chip Ros3 kernel (one persistent block per SM, OC_SLOTS cells in flight), its own
// translation unit (mechgen/onchip.py, ros3_onchip.inc).
#include "kpp_onchip.h"
#include "_gen/onchip_g.cuh"
namespace oc_g {
__constant__ double c_lit[NLIT];
}
#define OC_NS oc_g
#define OC_KERNEL ros3_onchip_g
#define OC_SLOTS 15
#define OC_W 1
#include "ros3_onchip.inc"
extern "C" const unsigned short mistra_oc_tables_g[];
extern "C" const size_t mistra_oc_tables_g_count;
namespace oc_g {
static cudaError_t launch(const KppBatch &b, int blocks, cudaStream_t st)
{
  ros3_onchip_g<<<blocks, NTHREADS, SMEM_BYTES, st>>>(b);
  return cudaGetLastError();
}
static cudaError_t set_lit(const double *h, cudaStream_t st)
{
  return cudaMemcpyToSymbolAsync(c_lit, h, sizeof(double) * NLIT, 0, cudaMemcpyHostToDevice, st);
}
}  // namespace oc_g

const KppOnchipInfo *kpp_onchip_info_g()
{
  using namespace oc_g;
  static const KppOnchipInfo info = {(const void *)ros3_onchip_g, launch, set_lit, mistra_oc_tables_g,
                                     mistra_oc_tables_g_count, coef_literals, NLIT, NTHREADS, SMEM_BYTES, T, SLOTS};
  return &info;
}
