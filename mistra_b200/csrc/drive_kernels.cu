// Gather / scatter halves of gas_drive / aer_drive / tot_drive on the device (include/mistra_drive.h):
// CUDA kernels + C-ABI entries.  Role in the reference: the copies of /root/reference/src/aer.f:146-178 and
// 233-245 (and their gas.f / tot.f counterparts) for one layer, here one CTA per batch cell: the map
// entries are spread over the threads (the KPP side of every copy is a contiguous row of var / fix, the
// model side a row of the layer), the liquid-phase rows are clamped by the same CTA before / after.
// HBM-bound copies, a few kB per cell.
#include "../../include/mistra_drive.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <string>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int DRV_THREADS = 256;

__device__ __forceinline__ double *row_of(const mistra_drive_args &a, int arr, long long k)
{
  switch (arr) {
    case 0: return a.s1 + k * a.j1;
    case 1: return a.s3 + k * a.j5;
    case 2: return a.sl1 + k * (long long)a.nkc * a.j2;
    default: return a.sion1 + k * (long long)a.nkc * a.j6;
  }
}

__global__ void __launch_bounds__(DRV_THREADS) drive_gather_kernel(long long ncell, mistra_drive_args a)
{
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    const long long k = a.layer[c];
    if (a.clamp_liquid) {                                  // aer.f:158-159
      double *l = a.sl1 + k * (long long)a.nkc * a.j2, *s = a.sion1 + k * (long long)a.nkc * a.j6;
      for (int q = threadIdx.x; q < a.nkc * a.j2; q += DRV_THREADS) l[q] = fmax(0.0, l[q]);
      for (int q = threadIdx.x; q < a.nkc * a.j6; q += DRV_THREADS) s[q] = fmax(0.0, s[q]);
      __syncthreads();
    }
    double *var = a.var + c * a.nvar, *fix = a.fix + c * a.nfix;
    for (int m = threadIdx.x; m < a.nmap; m += DRV_THREADS) {
      const int kp = a.map_kpp[m];
      const double v = row_of(a, a.map_arr[m], k)[a.map_off[m]];
      if (kp <= a.nvar) var[kp - 1] = v; else fix[kp - a.nvar - 1] = v;
    }
    if (threadIdx.x == 0) {                                // aer.f:153-170
      const double air = a.air[c];
      const double c21 = a.f32_literals ? (double)0.21f : 0.21, c79 = a.f32_literals ? (double)0.79f : 0.79;
      if (a.indf_o2) fix[a.indf_o2 - 1] = c21 * air;
      if (a.indf_h2o) fix[a.indf_h2o - 1] = a.h2o[c];
      if (a.indf_n2) fix[a.indf_n2 - 1] = c79 * air;
    } else if (threadIdx.x <= 4) {
      const int b = threadIdx.x - 1;
      if (a.indf_h2ol[b]) {
        const double cv = a.cvv[c * 4 + b];
        const double c55 = a.f32_literals ? (double)55.55f : 55.55;
        fix[a.indf_h2ol[b] - 1] = (cv > 0.0) ? c55 / cv : 0.0;
      }
    }
  }
}

__global__ void __launch_bounds__(DRV_THREADS) drive_scatter_kernel(long long ncell, mistra_drive_args a)
{
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    const long long k = a.layer[c];
    const double *var = a.var + c * a.nvar, *fix = a.fix + c * a.nfix;
    for (int m = threadIdx.x; m < a.nmap; m += DRV_THREADS) {     // aer.f:233-245
      const int kp = a.map_kpp[m];
      row_of(a, a.map_arr[m], k)[a.map_off[m]] = (kp <= a.nvar) ? var[kp - 1] : fix[kp - a.nvar - 1];
    }
    if (a.clip_negative) {                                        // kpp.f90:4472-4477
      __syncthreads();
      double *r1 = a.s1 + k * a.j1, *r3 = a.s3 + k * a.j5;
      double *l = a.sl1 + k * (long long)a.nkc * a.j2, *s = a.sion1 + k * (long long)a.nkc * a.j6;
      for (int q = threadIdx.x; q < a.j1; q += DRV_THREADS) if (r1[q] < 0.0) r1[q] = 0.0;
      for (int q = threadIdx.x; q < a.j5; q += DRV_THREADS) if (r3[q] < 0.0) r3[q] = 0.0;
      for (int q = threadIdx.x; q < a.nkc * a.j2; q += DRV_THREADS) l[q] = fmax(0.0, l[q]);
      for (int q = threadIdx.x; q < a.nkc * a.j6; q += DRV_THREADS) s[q] = fmax(0.0, s[q]);
    }
    __syncthreads();
  }
}

std::atomic<long long> g_launches{0};

int check(int64_t ncell, const mistra_drive_args *a, bool gather)
{
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null arguments");
  if (a->nvar < 1 || a->nfix < 0 || a->j1 < 0 || a->j5 < 0 || a->j2 < 0 || a->j6 < 0 || a->nkc < 1 || a->nkc > 4 ||
      a->nmap < 0)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes");
  if (a->nmap > 0 && (!a->map_kpp || !a->map_arr || !a->map_off))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null map");
  if (ncell > 0 && (!a->layer || !a->var || (a->nfix > 0 && !a->fix) || !a->s1 || !a->s3 || !a->sl1 || !a->sion1))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (gather && ncell > 0) {
    if (!a->air || !a->h2o) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null air / h2o");
    const int32_t f[7] = {a->indf_o2, a->indf_h2o, a->indf_n2, a->indf_h2ol[0], a->indf_h2ol[1], a->indf_h2ol[2],
                          a->indf_h2ol[3]};
    for (int i = 0; i < 7; ++i)
      if (f[i] < 0 || f[i] > a->nfix) return mistra_internal_fail(MISTRA_KPP_EINVAL, "FIX position out of range");
    if ((f[3] || f[4] || f[5] || f[6]) && !a->cvv) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null cvv");
  }
  return 0;
}

int launch(int64_t ncell, const mistra_drive_args *a, void *stream, bool gather)
{
  int rc = check(ncell, a, gather);
  if (rc) return rc;
  if (ncell == 0) return 0;
  int dev = -1, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess)
    return mistra_internal_fail(e == cudaErrorNoDevice ? MISTRA_KPP_ENODEVICE : MISTRA_KPP_ECUDA, cudaGetErrorString(e));
  long long blocks = (long long)sms * 8;
  if (blocks > ncell) blocks = ncell;
  if (gather) drive_gather_kernel<<<(int)blocks, DRV_THREADS, 0, (cudaStream_t)stream>>>(ncell, *a);
  else drive_scatter_kernel<<<(int)blocks, DRV_THREADS, 0, (cudaStream_t)stream>>>(ncell, *a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return mistra_internal_fail(MISTRA_KPP_ECUDA, cudaGetErrorString(e));
  g_launches.fetch_add(1);
  return 0;
}

}  // namespace

extern "C" {

int mistra_drive_gather_device(int64_t ncell, const mistra_drive_args *d_a, void *stream)
{
  return launch(ncell, d_a, stream, true);
}

int mistra_drive_scatter_device(int64_t ncell, const mistra_drive_args *d_a, void *stream)
{
  return launch(ncell, d_a, stream, false);
}

int64_t mistra_drive_launch_count(void) { return g_launches.load(); }

}  // extern "C"
