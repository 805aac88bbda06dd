// SUBROUTINE konc on the device (include/mistra_konc.h): CUDA kernel + C-ABI entries.
// Role in the reference: the layer loop of /root/reference/src/kpp.f90:3434-3588.
//
// Mapping: one CTA per layer.  The loop over dry classes is sequential (every class rescales
// what the previous ones left in the two bins it connects), but the species are independent:
// thread ia first works out the class' direction and fraction delta (kpp.f90:3438-3489), then
// thread l carries the four bin values of species l in registers through the 70 classes
// (3490-3500) and the final "too few droplets" transfer (3566-3586).  HBM-bound: 2*(4*(j2+j6))
// + 6*nka + 8 doubles per layer = 14.7 kB for the reference's sizes.  Compiled without FMA
// contraction (build.py); products and sums as written in the reference.
#include "../../include/mistra_konc.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int KONC_THREADS = 192;
constexpr int KONC_MAX_NKA = 256;

__global__ void __launch_bounds__(KONC_THREADS) konc_kernel(long long ncell, mistra_konc_args a)
{
  __shared__ double s_delta[KONC_MAX_NKA];   // fraction moved by class ia (0 = leave unchanged)
  __shared__ signed char s_dir[KONC_MAX_NKA]; // +1: aerosol bin loses (ii = 1|2), -1: droplet bin loses
  __shared__ int s_warn[3];
  const int nka = a.nka, ka = a.ka, j2 = a.j2, j6 = a.j6, nsp = j2 + j6;
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 3) s_warn[threadIdx.x] = 0;
    __syncthreads();
    for (int ia = threadIdx.x; ia < nka; ia += blockDim.x) {
      const long long o = c * nka + ia;
      const int ba = ia < ka ? 0 : 1;                       // aerosol bin of this class (0-based), droplet bin = ba + 2
      const double poa = a.part_o_a[o], pod = a.part_o_d[o];
      const double dp_a = poa - a.part_n_a[o];
      const double dp_d = pod - a.part_n_d[o];
      if (fabs(dp_a + dp_d) > 1.e-10) atomicAdd(&s_warn[0], 1);
      const bool from_a = dp_a >= 1.e-10;
      const double xs = (fabs(dp_a) < 1.e-10) ? 0.0 : 1.0;
      double delta;
      if (from_a) {
        const double v2 = a.vol2[c * 4 + ba];
        delta = (v2 > 0. && poa > 0.) ? a.vol1_a[o] / v2 * dp_a / poa * xs : 0.0;
      } else {
        const double v2 = a.vol2[c * 4 + ba + 2];
        delta = (v2 > 0. && pod > 0.) ? a.vol1_d[o] / v2 * dp_d / pod * xs : 0.0;
      }
      if (delta < 0.0) { atomicAdd(&s_warn[1], 1); delta = 0.0; }
      else if (delta > 1.0) { atomicAdd(&s_warn[2], 1); delta = 0.0; }
      s_delta[ia] = delta;
      s_dir[ia] = (delta > 0.0) ? (from_a ? 1 : -1) : 0;     // 0: class leaves the species unchanged
    }
    __syncthreads();
    for (int l = threadIdx.x; l < nsp; l += blockDim.x) {
      double *base = (l < j2) ? a.sl1 + (c * 4) * j2 + l : a.sion1 + (c * 4) * j6 + (l - j2);
      const int stride = (l < j2) ? j2 : j6;
      double s0 = base[0], s1 = base[stride], s2 = base[2 * stride], s3 = base[3 * stride];
      // small classes exchange between bins 1 <-> 3, large ones between 2 <-> 4 (two loops, integer tests only)
      const int kas = ka < nka ? ka : nka;
      for (int ia = 0; ia < kas; ++ia) {
        const int dir = s_dir[ia];
        if (dir == 0) continue;
        const double delta = s_delta[ia];
        if (dir > 0) { const double del = s0 * delta; s0 = fmax(0.0, s0 - del); s2 = fmax(0.0, s2 + del); }
        else         { const double del = s2 * delta; s2 = fmax(0.0, s2 - del); s0 = fmax(0.0, s0 + del); }
      }
      for (int ia = kas; ia < nka; ++ia) {
        const int dir = s_dir[ia];
        if (dir == 0) continue;
        const double delta = s_delta[ia];
        if (dir > 0) { const double del = s1 * delta; s1 = fmax(0.0, s1 - del); s3 = fmax(0.0, s3 + del); }
        else         { const double del = s3 * delta; s3 = fmax(0.0, s3 - del); s1 = fmax(0.0, s1 + del); }
      }
      if (a.pntot[c * 4 + 2] < 1.e-7) { s0 = s0 + fmax(0.0, s2); s2 = 0.0; }
      if (a.pntot[c * 4 + 3] < 1.e-7) { s1 = s1 + fmax(0.0, s3); s3 = 0.0; }
      base[0] = s0; base[stride] = s1; base[2 * stride] = s2; base[3 * stride] = s3;
    }
    if (a.warn && threadIdx.x < 3) a.warn[c * 3 + threadIdx.x] = s_warn[threadIdx.x];
  }
}

std::mutex g_mu;
std::atomic<long long> g_launches{0};
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_scratch[16];

#define CKC(call)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return mistra_internal_fail(e_ == cudaErrorMemoryAllocation ? MISTRA_KPP_ENOMEM   \
                                  : (e_ == cudaErrorNoDevice ? MISTRA_KPP_ENODEVICE     \
                                                             : MISTRA_KPP_ECUDA),       \
                                  std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

int check(int64_t ncell, const mistra_konc_args *a)
{
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null arguments");
  if (a->nka < 1 || a->nka > KONC_MAX_NKA || a->ka < 0 || a->ka > a->nka || a->j2 < 0 || a->j6 < 0 ||
      a->j2 + a->j6 < 1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (1 <= nka <= 256, 0 <= ka <= nka, j2 + j6 >= 1)");
  if (ncell > 0 && (!a->vol1_a || !a->vol1_d || !a->part_o_a || !a->part_o_d || !a->part_n_a || !a->part_n_d ||
                    !a->vol2 || !a->pntot || (a->j2 > 0 && !a->sl1) || (a->j6 > 0 && !a->sion1)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  return 0;
}

}  // namespace

extern "C" {

int mistra_konc_device(int64_t ncell, const mistra_konc_args *d_a, void *stream)
{
  int rc = check(ncell, d_a);
  if (rc) return rc;
  if (ncell == 0) return 0;
  int dev = -1, sms = 0;
  CKC(cudaGetDevice(&dev));
  CKC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = (long long)sms * 10;     // 10 CTAs of 192 threads per SM
  if (blocks > ncell) blocks = ncell;
  konc_kernel<<<(int)blocks, KONC_THREADS, 0, (cudaStream_t)stream>>>(ncell, *d_a);
  CKC(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mistra_konc(int64_t ncell, const mistra_konc_args *a, void *stream)
{
  int rc = check(ncell, a);
  if (rc) return rc;
  if (ncell == 0) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  CKC(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, nka = a->nka;
  struct Item { const void *h; size_t bytes; bool out; void **slot; };
  mistra_konc_args d = *a;
  std::vector<Item> items = {
      {a->vol1_a, n * nka * 8, false, (void **)&d.vol1_a}, {a->vol1_d, n * nka * 8, false, (void **)&d.vol1_d},
      {a->part_o_a, n * nka * 8, false, (void **)&d.part_o_a}, {a->part_o_d, n * nka * 8, false, (void **)&d.part_o_d},
      {a->part_n_a, n * nka * 8, false, (void **)&d.part_n_a}, {a->part_n_d, n * nka * 8, false, (void **)&d.part_n_d},
      {a->vol2, n * 32, false, (void **)&d.vol2}, {a->pntot, n * 32, false, (void **)&d.pntot},
      {a->sl1, n * 4 * a->j2 * 8, true, (void **)&d.sl1}, {a->sion1, n * 4 * a->j6 * 8, true, (void **)&d.sion1},
      {a->warn, n * 12, true, (void **)&d.warn}};
  size_t total = 0;
  for (auto &it : items)
    if (it.h && it.bytes) total += (it.bytes + 255) & ~(size_t)255;
  Scratch &sc = g_scratch[dev];
  if (sc.bytes < total) {
    if (sc.p) { CKC(cudaDeviceSynchronize()); cudaFree(sc.p); sc.p = nullptr; sc.bytes = 0; }
    CKC(cudaMalloc(&sc.p, total));
    sc.bytes = total;
  }
  char *p = sc.p;
  for (auto &it : items) {
    if (!it.h || !it.bytes) { *it.slot = nullptr; continue; }
    *it.slot = p;
    if (it.h != (const void *)a->warn) CKC(cudaMemcpyAsync(p, it.h, it.bytes, cudaMemcpyHostToDevice, st));
    p += (it.bytes + 255) & ~(size_t)255;
  }
  if ((rc = mistra_konc_device(ncell, &d, stream))) return rc;
  for (auto &it : items)
    if (it.h && it.bytes && it.out) CKC(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKC(cudaStreamSynchronize(st));
  return 0;
}

int64_t mistra_konc_launch_count(void) { return g_launches.load(); }

}  // extern "C"
