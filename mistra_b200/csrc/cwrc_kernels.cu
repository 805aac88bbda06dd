// SUBROUTINE cw_rc on the device (include/mistra_cwrc.h): CUDA kernel + C-ABI entries.
// Role in the reference: the layer loop of /root/reference/src/kpp.f90:2260-2412.
//
// Mapping: one CTA (256 threads) per layer; the layer's 70 x 70 spectrum goes to shared memory with
// 16-byte asynchronous copies (the whole tile in flight at once), then one warp per dry class at a
// time: the lanes stride over the class' water bins (rq and e come from the L1/L2-resident grid
// tables) and keep, over all classes of their warp, the sums volume cw, volume x radius rc and water
// mass cm (kpp.f90:2285-2322) of the four chemistry bins (aerosol part jt <= kw / droplet part, small
// classes ia <= ka / large ones); one fixed butterfly per sum adds the lanes, twelve threads add the
// warps in order and four apply the switches (2335-2410).  HBM-bound: nka*nkt*8 = 39.2 kB read per
// layer, 128 B written.  No FMA contraction (build.py).
#include "tma_bulk.h"
#include "../../include/mistra_cwrc.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int CWRC_THREADS = 256;
constexpr int CWRC_MAX = 128;   // nka, nkt

__device__ __forceinline__ double warp_sum(double v)
{
  // fixed butterfly: the same tree for every class and layer, so the result is deterministic
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__global__ void __launch_bounds__(CWRC_THREADS) cwrc_kernel(long long ncell, mistra_cwrc_args a)
{
  __shared__ double s_p[(CWRC_THREADS / 32) * 12];   // per warp: [bin][cw, rc, cm]
  __shared__ double s_b[12];                 // sums of the chemistry bins
  const int nka = a.nka, nkt = a.nkt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double xpi = 4.0 / 3.0 * 3.1415926535897932;   // kpp.f90:2204, constants.f90 pi
  extern __shared__ __align__(16) double s_ff[];       // [nka][nkt] the layer's spectrum
  const int ntile = nka * nkt;
  __shared__ unsigned long long s_bar;       // completion barrier of the tile's bulk copy
  if (threadIdx.x == 0) tma::mbar_init(&s_bar, 1);
  unsigned phase = 0;
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    __syncthreads();
    const double *gf = a.ff + (size_t)c * ntile;
    if (tma::bulk_ok(gf, (size_t)ntile * 8)) {
      // the layer's 70 x 70 spectrum is one contiguous 39.2 kB block: ONE bulk (TMA) copy, completion on the mbarrier
      if (threadIdx.x == 0) tma::bulk_load(s_ff, gf, (unsigned)(ntile * 8), &s_bar);
      tma::mbar_wait(&s_bar, phase);
      phase ^= 1u;
    } else {
      // unaligned caller buffer: 16- / 8-byte asynchronous copies (no register staging)
      if ((ntile & 1) == 0 && ((reinterpret_cast<unsigned long long>(gf) & 15ull) == 0ull)) {
        for (int q = threadIdx.x; q < (ntile >> 1); q += blockDim.x)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(s_ff + 2 * q)),
                       "l"(gf + 2 * q) : "memory");
      } else {
        for (int q = threadIdx.x; q < ntile; q += blockDim.x)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(s_ff + q)),
                       "l"(gf + q) : "memory");
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
      __syncthreads();
    }
    const double *f = s_ff;
    // one warp per dry class at a time: lanes stride over the water bins (coalesced loads of ff and
    // rq); each lane keeps, over all classes of its warp, the sums of the four chemistry bins
    double acc[12];                          // [bin][cw, rc, cm], compile-time indices only
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[q] = 0.0;
    for (int ia = warp + a.ial - 1; ia < nka; ia += nwarp) {
      const int kwa = a.kw[ia];
      const bool small = ia < a.ka;          // warp-uniform: bins 1|3, else 2|4
      for (int jt = lane; jt < nkt; jt += 32) {
        const double ffv = f[ia * nkt + jt], r = a.rq[ia * nkt + jt];
        const double x0 = ffv * xpi * (r * r * r);
        const double x1 = x0 * r, x2 = ffv * a.e[jt];
        const bool aer = jt < kwa;
        if (small) {
          if (aer) { acc[0] = acc[0] + x0; acc[1] = acc[1] + x1; acc[2] = acc[2] + x2; }
          else     { acc[6] = acc[6] + x0; acc[7] = acc[7] + x1; acc[8] = acc[8] + x2; }
        } else {
          if (aer) { acc[3] = acc[3] + x0; acc[4] = acc[4] + x1; acc[5] = acc[5] + x2; }
          else     { acc[9] = acc[9] + x0; acc[10] = acc[10] + x1; acc[11] = acc[11] + x2; }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[q] = warp_sum(acc[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 12; ++q) s_p[warp * 12 + q] = acc[q];
    }
    __syncthreads();
    // twelve threads: (bin, quantity) over the warps, in warp order
    if (threadIdx.x < 12) {
      double t = 0.0;
      for (int w = 0; w < nwarp; ++w) t = t + s_p[w * 12 + threadIdx.x];
      s_b[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      const int kc = threadIdx.x;
      const double cw = s_b[kc * 3], rc = s_b[kc * 3 + 1], cm = s_b[kc * 3 + 2];
      const double feu = a.feu[c];
      a.rc[c * 4 + kc] = (cw > 0.0) ? rc / cw * 1.e-6 : 0.0;
      a.cw[c * 4 + kc] = cw * 1.e-12;
      double cmo = 0.0, cv = 0.0;
      if (!(feu < fmin(a.xcryssulf, a.xcrysss))) {
        bool on;
        if (kc == 0) on = (cw >= 1.e-1) && ((a.cloud[c * 4 + 0] && feu >= a.xcryssulf) || (feu >= a.xdelisulf));
        else if (kc == 1) on = (cw >= 1.e-1) && ((a.cloud[c * 4 + 1] && feu >= a.xcrysss) || (feu >= a.xdeliss));
        else on = (cw >= 1.e2);
        if (on) { cmo = cm * 1.e-3; cv = 1.e9 / cw; }
      }
      a.cm[c * 4 + kc] = cmo;
      a.conv2[c * 4 + kc] = cv;
    }
  }
}

std::mutex g_mu;
std::atomic<long long> g_launches{0};
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_scratch[16];
bool g_attr[16] = {};

#define CKW(call)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return mistra_internal_fail(e_ == cudaErrorMemoryAllocation ? MISTRA_KPP_ENOMEM   \
                                  : (e_ == cudaErrorNoDevice ? MISTRA_KPP_ENODEVICE     \
                                                             : MISTRA_KPP_ECUDA),       \
                                  std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

int check(int64_t ncell, const mistra_cwrc_args *a)
{
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null arguments");
  if (a->nka < 1 || a->nka > CWRC_MAX || a->nkt < 1 || a->nkt > CWRC_MAX || a->ka < 0 || a->ka > a->nka ||
      a->ial < 1 || a->ial > 2)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (nka, nkt <= 128, 0 <= ka <= nka, ial = 1|2)");
  if (!a->kw || !a->e || !a->rq) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null grid array");
  if (ncell > 0 && (!a->ff || !a->feu || !a->cloud || !a->rc || !a->cw || !a->cm || !a->conv2))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  return 0;
}

}  // namespace

extern "C" {

int mistra_cwrc_device(int64_t ncell, const mistra_cwrc_args *d_a, void *stream)
{
  int rc = check(ncell, d_a);
  if (rc) return rc;
  if (ncell == 0) return 0;
  int dev = -1, sms = 0;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  CKW(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(double) * (size_t)d_a->nka * d_a->nkt;
  if (!g_attr[dev]) {
    CKW(cudaFuncSetAttribute(cwrc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
    g_attr[dev] = true;
  }
  int per_sm = (int)((224 * 1024) / (smem + 2048));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long blocks = (long long)sms * per_sm;
  if (blocks > ncell) blocks = ncell;
  cwrc_kernel<<<(int)blocks, CWRC_THREADS, smem, (cudaStream_t)stream>>>(ncell, *d_a);
  CKW(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mistra_cwrc(int64_t ncell, const mistra_cwrc_args *a, void *stream)
{
  int rc = check(ncell, a);
  if (rc) return rc;
  if (ncell == 0) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, nka = a->nka, nkt = a->nkt;
  struct Item { const void *h; size_t bytes; bool in, out; void **slot; };
  mistra_cwrc_args d = *a;
  std::vector<Item> items = {
      {a->kw, nka * 4, true, false, (void **)&d.kw}, {a->e, nkt * 8, true, false, (void **)&d.e},
      {a->rq, nka * nkt * 8, true, false, (void **)&d.rq}, {a->ff, n * nka * nkt * 8, true, false, (void **)&d.ff},
      {a->feu, n * 8, true, false, (void **)&d.feu}, {a->cloud, n * 16, true, false, (void **)&d.cloud},
      {a->rc, n * 32, false, true, (void **)&d.rc}, {a->cw, n * 32, false, true, (void **)&d.cw},
      {a->cm, n * 32, false, true, (void **)&d.cm}, {a->conv2, n * 32, false, true, (void **)&d.conv2}};
  size_t total = 0;
  for (auto &it : items) total += (it.bytes + 255) & ~(size_t)255;
  Scratch &sc = g_scratch[dev];
  if (sc.bytes < total) {
    if (sc.p) { CKW(cudaDeviceSynchronize()); cudaFree(sc.p); sc.p = nullptr; sc.bytes = 0; }
    CKW(cudaMalloc(&sc.p, total));
    sc.bytes = total;
  }
  char *p = sc.p;
  for (auto &it : items) {
    *it.slot = p;
    if (it.in) CKW(cudaMemcpyAsync(p, it.h, it.bytes, cudaMemcpyHostToDevice, st));
    p += (it.bytes + 255) & ~(size_t)255;
  }
  if ((rc = mistra_cwrc_device(ncell, &d, stream))) return rc;
  for (auto &it : items)
    if (it.out) CKW(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKW(cudaStreamSynchronize(st));
  return 0;
}

int64_t mistra_cwrc_launch_count(void) { return g_launches.load(); }

}  // extern "C"
