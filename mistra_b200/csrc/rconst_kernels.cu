// Update_RCONST_x on the device (include/mistra_rconst_cuda.h): one thread per cell
// evaluates the NREACT rate-constant expressions of its mechanism - the same generated
// right-hand sides (_gen/rconst_<x>.inc) and the same rate-law functions (rate_laws.h) the
// host library compiles.  Role in the reference: Update_RCONST_g/_a/_t
// (/root/reference/src/gas.f:275-666, aer.f:304-1400, tot.f:1040-2805) with the rate laws of
// /root/reference/src/kpp.f90:7127-8373.  Compiled without FMA contraction (build.py).
#include "../../include/mistra_rconst_cuda.h"
#include "../../include/mistra_kpp.h"
#include "rconst_common.h"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

using rconst_common::kDims;

__device__ void rc_g(const rate_ctx *cx, double *RC)
{
#include "_gen/rconst_g.inc"
}
__device__ void rc_a(const rate_ctx *cx, double *RC)
{
#include "_gen/rconst_a.inc"
}
__device__ void rc_t(const rate_ctx *cx, double *RC)
{
#include "_gen/rconst_t.inc"
}

struct RcIn {
  long long ncell;
  int nvar, nreact, nkc, nspec;
  const double *cb1, *scal, *ph_rat, *conc, *yhenry, *yxkmt, *ykef, *ykeb, *yxkmtd, *yxeq, *ycw, *ycwd;
  const double *zeros;   // nspec*4 zeros for absent arrays
  rate_ctx proto;        // literal switch and species indices
};

template <int MECH>
__global__ void __launch_bounds__(128) rconst_kernel(RcIn in, double *__restrict__ rconst)
{
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= in.ncell) return;
  rate_ctx cx = in.proto;
  const double *cb = in.cb1 + 4 * c, *sc = in.scal + 13 * c;
  cx.aircc = cb[0]; cx.te = cb[1]; cx.h2oppm = cb[2]; cx.pk = cb[3];
  cx.conv1 = sc[0]; cx.xhal = sc[1]; cx.xiod = sc[2]; cx.xhet1 = sc[3]; cx.xhet2 = sc[4];
  for (int k = 0; k < 4; ++k) { cx.xliq[k] = sc[5 + k]; cx.cvv[k] = sc[9 + k]; }
  cx.ph_rat = in.ph_rat + (size_t)MISTRA_NPHRXN * c;
  cx.C = in.conc + (size_t)in.nspec * c;
  cx.FIX = cx.C + in.nvar;
  cx.yhenry = in.yhenry ? in.yhenry + (size_t)in.nspec * c : in.zeros;
  cx.yxkmt = in.yxkmt ? in.yxkmt + (size_t)in.nspec * in.nkc * c : in.zeros;
  cx.ykef = in.ykef ? in.ykef + (size_t)in.nspec * in.nkc * c : in.zeros;
  cx.ykeb = in.ykeb ? in.ykeb + (size_t)in.nspec * in.nkc * c : in.zeros;
  cx.yxkmtd = in.yxkmtd ? in.yxkmtd + (size_t)in.nspec * 2 * c : in.zeros;
  cx.yxeq = in.yxeq ? in.yxeq + (size_t)in.nspec * c : in.zeros;
  cx.ycw = in.ycw ? in.ycw + (size_t)in.nkc * c : in.zeros;
  cx.ycwd = in.ycwd ? in.ycwd + (size_t)2 * c : in.zeros;
  double *RC = rconst + (size_t)in.nreact * c;
  if (MECH == 0) rc_g(&cx, RC);
  else if (MECH == 1) rc_a(&cx, RC);
  else rc_t(&cx, RC);
}

std::mutex g_mu;
std::atomic<long long> g_launches{0};
double *g_zeros[16] = {};

}  // namespace

extern "C" {

int mistra_rconst_update_device(int mech, const mistra_rate_inputs *in, double *d_rconst, void *stream)
{
  if (mech < 0 || mech > 2 || !in || in->ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad mechanism / inputs");
  if (in->ncell == 0) return 0;
  if (!d_rconst || !in->cb1 || !in->scal || !in->ph_rat || !in->conc)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null rconst / cb1 / scal / ph_rat / conc");
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 16)
    return mistra_internal_fail(MISTRA_KPP_ENODEVICE, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
  cudaStream_t st = (cudaStream_t)stream;
  const rconst_common::MechDims dm = kDims[mech];
  if (!g_zeros[dev]) {
    const size_t nb = sizeof(double) * 424 * 4;
    if ((e = cudaMalloc(&g_zeros[dev], nb)) != cudaSuccess || (e = cudaMemset(g_zeros[dev], 0, nb)) != cudaSuccess)
      return mistra_internal_fail(MISTRA_KPP_ECUDA, std::string("zeros: ") + cudaGetErrorString(e));
  }
  RcIn k;
  memset(&k, 0, sizeof k);
  k.ncell = in->ncell;
  k.nvar = dm.nvar; k.nreact = dm.nreact; k.nkc = dm.nkc; k.nspec = dm.nvar + dm.nfix;
  k.cb1 = in->cb1; k.scal = in->scal; k.ph_rat = in->ph_rat; k.conc = in->conc;
  k.yhenry = in->yhenry; k.yxkmt = in->yxkmt; k.ykef = in->ykef; k.ykeb = in->ykeb;
  k.yxkmtd = in->yxkmtd; k.yxeq = in->yxeq; k.ycw = in->ycw; k.ycwd = in->ycwd;
  k.zeros = g_zeros[dev];
  k.proto.nspec = k.nspec;
  k.proto.f32 = in->f32_literals ? 1 : 0;
  rconst_common::fill_indices(mech, &k.proto);
  const int threads = 128;
  const long long blocks = (in->ncell + threads - 1) / threads;
  if (mech == 0) rconst_kernel<0><<<(unsigned)blocks, threads, 0, st>>>(k, d_rconst);
  else if (mech == 1) rconst_kernel<1><<<(unsigned)blocks, threads, 0, st>>>(k, d_rconst);
  else rconst_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(k, d_rconst);
  if ((e = cudaGetLastError()) != cudaSuccess)
    return mistra_internal_fail(MISTRA_KPP_ECUDA, std::string("rconst_kernel: ") + cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return 0;
}

int64_t mistra_rconst_launch_count(void) { return g_launches.load(); }

}  // extern "C"
