// SUBROUTINE cw_rc on the device (include/mistra_cwrc.h): CUDA kernel + C-ABI entries.
// Role in the reference: the layer loop of /root/reference/src/kpp.f90:2260-2412.
//
// Mapping: one CTA (256 threads) per layer.  The 70 x 70 spectrum is staged in shared memory
// with coalesced loads (rows padded to an odd stride), two threads per dry class then sum the
// class' aerosol part (jt <= kw) and droplet part in bin order - three sums each: volume cw,
// volume x radius rc, water mass cm (kpp.f90:2285-2322) - and four threads add the classes of
// their chemistry bin in class order and apply the switches (2335-2410).  HBM-bound: nka*nkt*8 =
// 39.2 kB read per layer, 128 B written.  No FMA contraction (build.py).
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

__global__ void __launch_bounds__(CWRC_THREADS) cwrc_kernel(long long ncell, mistra_cwrc_args a)
{
  extern __shared__ double sm[];
  const int nka = a.nka, nkt = a.nkt, ld = nkt | 1;
  double *s_ff = sm;                     // [nka][ld]
  double *s_p = s_ff + nka * ld;         // [2*nka][3] partial sums of (class, part)
  const double xpi = 4.0 / 3.0 * 3.1415926535897932;   // kpp.f90:2204, constants.f90 pi
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    __syncthreads();
    const double *f = a.ff + (size_t)c * nka * nkt;
    for (int q = threadIdx.x; q < nka * nkt; q += blockDim.x) {
      const int ia = q / nkt, jt = q - ia * nkt;
      s_ff[ia * ld + jt] = f[q];
    }
    __syncthreads();
    if (threadIdx.x < 2 * nka) {
      const int ia = threadIdx.x >> 1, part = threadIdx.x & 1, kwa = a.kw[ia];
      const int j0 = part ? kwa : 0, j1 = part ? nkt : kwa;
      double cw = 0.0, rc = 0.0, cm = 0.0;
      for (int jt = j0; jt < j1; ++jt) {
        const double ffv = s_ff[ia * ld + jt], r = a.rq[ia * nkt + jt];
        const double x0 = ffv * xpi * (r * r * r);
        cw = cw + x0;
        rc = rc + x0 * r;
        cm = cm + ffv * a.e[jt];
      }
      double *o = s_p + threadIdx.x * 3;
      o[0] = cw; o[1] = rc; o[2] = cm;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      const int kc = threadIdx.x, part = kc >> 1;
      const int lo = (kc & 1) ? a.ka : a.ial - 1, hi = (kc & 1) ? nka : a.ka;
      double cw = 0.0, rc = 0.0, cm = 0.0;
      for (int ia = lo; ia < hi; ++ia) {
        const double *p = s_p + (2 * ia + part) * 3;
        cw = cw + p[0]; rc = rc + p[1]; cm = cm + p[2];
      }
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
  const size_t smem = sizeof(double) * ((size_t)d_a->nka * (d_a->nkt | 1) + 6 * (size_t)d_a->nka);
  if (!g_attr[dev]) {
    CKW(cudaFuncSetAttribute(cwrc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    g_attr[dev] = true;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
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
