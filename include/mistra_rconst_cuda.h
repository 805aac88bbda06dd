/* mistra_rconst_cuda.h - Update_RCONST_x on the device (part of libmistra_kpp.so).
 *
 * "Next" row N1 of the hot-path scope: the rate constants of every cell of a
 * batch computed by a CUDA kernel from the same inputs Update_RCONST_g/_a/_t read
 * (/root/reference/src/gas.f:275-666, aer.f:304-1400, tot.f:1040-2805; rate laws
 * /root/reference/src/kpp.f90:7127-8373), so that RCONST never crosses PCIe and the serial
 * host loop over layers (gas.f:172, aer.f:216, tot.f:603) disappears.  The result feeds
 * mistra_kpp_integrate_device directly.
 *
 * `in` is a HOST struct (include/mistra_rconst.h) whose array members are DEVICE pointers
 * on the current device; NULL arrays mean all zero, exactly as for the host version.
 * d_rconst [ncell][NREACT] out.  Same expressions, same literal convention (f32_literals)
 * as mistra_rconst_update; exp/pow/log10 come from the CUDA math library, so the two agree
 * to a few ulp per operation (tests: 1e-12 relative), not to the bit.
 * Asynchronous on `stream` (NULL = legacy default stream).  Returns 0 or MISTRA_KPP_E*. */
#ifndef MISTRA_RCONST_CUDA_H
#define MISTRA_RCONST_CUDA_H
#include "mistra_rconst.h"
#ifdef __cplusplus
extern "C" {
#endif

int mistra_rconst_update_device(int mech, const mistra_rate_inputs *in, double *d_rconst,
                                void *stream);

/* Kernels launched by this entry since load. */
int64_t mistra_rconst_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
