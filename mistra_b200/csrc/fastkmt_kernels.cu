// SUBROUTINE fast_k_mt_a / fast_k_mt_t on the device (include/mistra_fastkmt.h): CUDA kernel + C-ABI
// entries.  Role in the reference: the layer loop of /root/reference/src/kpp.f90:2820-2945 (and
// 2558-2674), with FUNCTION vterm (str.f90:2793-2864).
//
// Mapping: one persistent CTA (512 threads) per SM, a layer at a time.  The grid tables - radius rqm
// and the chemistry bin of every grid point (aerosol / droplet part by kw, small / large classes by
// ka, 255 = not summed) - are put in shared memory once; per layer the 70 x 70 spectrum arrives with
// 16-byte asynchronous copies.  Pass 1 (a thread per grid point): q = rqm / freep into shared memory
// and the sedimentation sums (vterm only where ff != 0).  Pass 2: lanes = exchanged species (l and
// l + 32), warps stride over the grid points, two points per iteration so that every thread has four
// independent FP64 divisions in flight; each thread keeps its species' sums of the four bins in
// registers.  Points with ff == 0 and points of a bin without chemistry (cm <= 0) add exactly +0 in
// the reference and are skipped.  Partial sums are combined over the warps in warp order.
// FP64-bound: (nx + 1) divisions per populated grid point against 39.2 kB read per layer.
// No FMA contraction (build.py).
#include "../../include/mistra_fastkmt.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int FK_THREADS = 512;
constexpr int FK_WARPS = FK_THREADS / 32;
constexpr int FK_MAX = 128;   // nka, nkt
constexpr int FK_NKC = 4;     // chemistry bins of the 2-D spectrum

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// FUNCTION vterm(a,t,p), str.f90:2793-2864: Stokes with slip correction below 10 um, Beard's
// Best-number polynomial above (the reference has no third regime).
__device__ __forceinline__ double vterm(double a, double t, double p)
{
  const double g = 9.80665, gas_const = 8.3144743, M_air = 28.96546e-3;   // constants.f90
  const double r0 = gas_const / M_air, rhow = 1000.0;
  const double b0 = -.318657e+1, b1 = .992696e+0, b2 = -.153193e-2, b3 = -.987059e-3, b4 = -.578878e-3,
               b5 = +.855176e-4, b6 = -.327815e-5;
  const double c1 = 2.0 * g / 9.0, c2 = 1.26, P0 = 101325, T0 = 293.15, lambda0 = 6.6e-8;
  const double c3 = c2 * lambda0 * P0 / T0, c4 = 32.0 * g / 3.0;
  const double rho_a = p / (r0 * t);
  const double eta = 3.7957e-06 + 4.9e-08 * t;
  if (a <= 1.e-5) return c1 * a * a * (rhow - rho_a) / eta * (1.0 + c3 * t / (a * p));
  const double best = c4 * (a * a * a) * (rhow - rho_a) * rho_a / (eta * eta);
  const double x = log(best);
  double y = b6 * x + b5;
  y = y * x + b4;
  y = y * x + b3;
  y = y * x + b2;
  y = y * x + b1;
  y = y * x + b0;
  return eta * exp(y) / (2. * rho_a * a);
}

#define FK_ACC(kc, v0, v1)                                                  \
  do {                                                                      \
    switch (kc) {                                                           \
      case 0: a00 = a00 + (v0); a10 = a10 + (v1); break;                    \
      case 1: a01 = a01 + (v0); a11 = a11 + (v1); break;                    \
      case 2: a02 = a02 + (v0); a12 = a12 + (v1); break;                    \
      default: a03 = a03 + (v0); a13 = a13 + (v1); break;                   \
    }                                                                       \
  } while (0)

__global__ void __launch_bounds__(FK_THREADS, 1) fastkmt_kernel(long long ncell, mistra_fastkmt_args a)
{
  extern __shared__ __align__(16) double smem[];
  const int nka = a.nka, nkt = a.nkt, ntile = nka * nkt, npad = (ntile + 1) & ~1;
  double *s_f = smem;                       // [ntile] the layer's spectrum
  double *s_r = s_f + npad;                 // [ntile] rqm = rq * 1e-6 (kpp.f90:2806)
  double *s_q = s_r + npad;                 // [ntile] rqm / freep(k)
  double *s_red = s_q + npad;               // [FK_WARPS][FK_NKC][64] partial sums of pass 2
  double *s_vt = s_red + FK_WARPS * FK_NKC * 64;   // [FK_WARPS][FK_NKC]
  unsigned char *s_kc = (unsigned char *)(s_vt + FK_WARPS * FK_NKC);   // [ntile] bin of the point, 255 = none
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double z4pi3 = 4.0 * 3.1415926535897932 / 3.0;   // kpp.f90:2758
  const int nx = a.nx, nspec = a.nspec, nkc = a.nkc;

  for (int q = threadIdx.x; q < ntile; q += FK_THREADS) {
    const int ia = q / nkt, jt = q - ia * nkt;
    s_r[q] = a.rq[q] * 1.e-6;
    int kc = (jt < a.kw[ia] ? 0 : 2) + (ia < a.ka ? 0 : 1);          // kpp.f90:2869-2901
    if (kc >= a.nkc_l || (ia < a.ial - 1 && ia < a.ka)) kc = 255;    // ifeed == 2 drops the first small class
    s_kc[q] = (unsigned char)kc;
  }
  const int l0 = lane, l1 = lane + 32;
  const int sp0 = l0 < nx ? a.lex[l0] - 1 : -1, sp1 = l1 < nx ? a.lex[l1] - 1 : -1;

  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    __syncthreads();
    const double *gf = a.ff + (size_t)c * ntile;
    if ((ntile & 1) == 0) {
      for (int q = threadIdx.x; q < (ntile >> 1); q += FK_THREADS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(s_f + 2 * q)),
                     "l"(gf + 2 * q) : "memory");
    } else {
      for (int q = threadIdx.x; q < ntile; q += FK_THREADS)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(s_f + q)),
                     "l"(gf + q) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    // layer scalars while the tile is in flight
    const double freep = a.freep[c], tk = a.t[c], pk = a.p[c];
    unsigned onmask = 0;                     // bins with chemistry (cm > 0, kpp.f90:2826)
    for (int kc = 0; kc < a.nkc_l; ++kc)
      if (a.cm[c * nkc + kc] > 0.0) onmask |= 1u << kc;
    double vm0 = 0.0, vm1 = 0.0, x10 = 1.0, x11 = 1.0;     // idle lanes: 0 / (q + 1)
    if (sp0 >= 0) {
      const double al = a.alpha[c * nspec + sp0];
      vm0 = a.vmean[c * nspec + sp0];
      x10 = (al > 0.0) ? 4. / (3. * al) : 0.0;             // kpp.f90:2890
    }
    if (sp1 >= 0) {
      const double al = a.alpha[c * nspec + sp1];
      vm1 = a.vmean[c * nspec + sp1];
      x11 = (al > 0.0) ? 4. / (3. * al) : 0.0;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();

    // pass 1: q = rqm / freep, sedimentation sums (kpp.f90:2924-2927)
    {
      double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
      for (int q = threadIdx.x; q < ntile; q += FK_THREADS) {
        const double rqq = s_r[q], f = s_f[q];
        s_q[q] = rqq / freep;
        const int kc = s_kc[q];
        if (kc != 255 && f != 0.0) {
          const double xvs = vterm(rqq, tk, pk);
          const double term = rqq * rqq * rqq * xvs * f * 1.e6;
          if (kc == 0) v0 = v0 + term; else if (kc == 1) v1 = v1 + term; else if (kc == 2) v2 = v2 + term; else v3 = v3 + term;
        }
      }
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
      if (lane == 0) { s_vt[warp * 4 + 0] = v0; s_vt[warp * 4 + 1] = v1; s_vt[warp * 4 + 2] = v2; s_vt[warp * 4 + 3] = v3; }
    }
    __syncthreads();

    // pass 2: transfer coefficients (kpp.f90:2913-2922); warp-uniform control flow
    double a00 = 0.0, a01 = 0.0, a02 = 0.0, a03 = 0.0, a10 = 0.0, a11 = 0.0, a12 = 0.0, a13 = 0.0;
    if (onmask) {
      for (int q = warp; q < ntile; q += 2 * FK_WARPS) {
        const int qb = q + FK_WARPS;
        const int kcA = s_kc[q], kcB = qb < ntile ? s_kc[qb] : 255;
        const double fA = s_f[q], fB = qb < ntile ? s_f[qb] : 0.0;
        const bool actA = kcA != 255 && ((onmask >> kcA) & 1u) && fA != 0.0;
        const bool actB = kcB != 255 && ((onmask >> kcB) & 1u) && fB != 0.0;
        if (actA && actB) {
          const double rA = s_r[q], qA = s_q[q], rB = s_r[qb], qB = s_q[qb];
          const double xA0 = vm0 / (qA + x10), xA1 = vm1 / (qA + x11);
          const double xB0 = vm0 / (qB + x10), xB1 = vm1 / (qB + x11);
          const double tA0 = xA0 * rA * rA * fA * 1.e6, tA1 = xA1 * rA * rA * fA * 1.e6;
          const double tB0 = xB0 * rB * rB * fB * 1.e6, tB1 = xB1 * rB * rB * fB * 1.e6;
          FK_ACC(kcA, tA0, tA1);
          FK_ACC(kcB, tB0, tB1);
        } else if (actA) {
          const double rA = s_r[q], qA = s_q[q];
          const double xA0 = vm0 / (qA + x10), xA1 = vm1 / (qA + x11);
          const double tA0 = xA0 * rA * rA * fA * 1.e6, tA1 = xA1 * rA * rA * fA * 1.e6;
          FK_ACC(kcA, tA0, tA1);
        } else if (actB) {
          const double rB = s_r[qb], qB = s_q[qb];
          const double xB0 = vm0 / (qB + x10), xB1 = vm1 / (qB + x11);
          const double tB0 = xB0 * rB * rB * fB * 1.e6, tB1 = xB1 * rB * rB * fB * 1.e6;
          FK_ACC(kcB, tB0, tB1);
        }
      }
      double *r = s_red + warp * (FK_NKC * 64);
      r[0 * 64 + l0] = a00; r[1 * 64 + l0] = a01; r[2 * 64 + l0] = a02; r[3 * 64 + l0] = a03;
      r[0 * 64 + l1] = a10; r[1 * 64 + l1] = a11; r[2 * 64 + l1] = a12; r[3 * 64 + l1] = a13;
    }
    __syncthreads();

    // results (kpp.f90:2931-2937): warps in order
    if (threadIdx.x < FK_NKC * 64) {
      const int kc = threadIdx.x >> 6, l = threadIdx.x & 63;
      if (kc < a.nkc_l && l < nx && ((onmask >> kc) & 1u)) {
        const double cw = a.cw[c * nkc + kc];
        if (cw > 0.0) {
          double s = 0.0;
          for (int w = 0; w < FK_WARPS; ++w) s = s + s_red[w * (FK_NKC * 64) + kc * 64 + l];
          a.xkmt[((size_t)c * nkc + kc) * nspec + a.lex[l] - 1] = z4pi3 / cw * s;
        }
      }
    } else if (threadIdx.x < FK_NKC * 64 + FK_NKC) {
      const int kc = threadIdx.x - FK_NKC * 64;
      if (kc < a.nkc_l) {
        const double cw = a.cw[c * nkc + kc];
        if (cw > 0.0) {
          double s = 0.0;
          for (int w = 0; w < FK_WARPS; ++w) s = s + s_vt[w * 4 + kc];
          a.vt[c * nkc + kc] = z4pi3 / cw * s;
        }
      }
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

int check(int64_t ncell, const mistra_fastkmt_args *a)
{
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null arguments");
  if (a->nka < 1 || a->nka > FK_MAX || a->nkt < 1 || a->nkt > FK_MAX || a->ka < 0 || a->ka > a->nka ||
      a->ial < 1 || a->ial > 2)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (nka, nkt <= 128, 0 <= ka <= nka, ial = 1|2)");
  if (a->nkc < 1 || a->nkc > FK_NKC || a->nkc_l < 0 || a->nkc_l > a->nkc)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad bin counts (1 <= nkc <= 4, 0 <= nkc_l <= nkc)");
  if (a->nx < 1 || a->nx > MISTRA_FASTKMT_MAXNX || a->nspec < 1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad species counts (1 <= nx <= 64, nspec >= 1)");
  if (!a->kw || !a->rq || !a->lex) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null grid / species table");
  if (ncell > 0 && (!a->ff || !a->freep || !a->t || !a->p || !a->cw || !a->cm || !a->alpha || !a->vmean ||
                    !a->xkmt || !a->vt))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  return 0;
}

size_t smem_bytes(const mistra_fastkmt_args *a)
{
  const size_t ntile = (size_t)a->nka * a->nkt, npad = (ntile + 1) & ~(size_t)1;
  return sizeof(double) * (3 * npad + FK_WARPS * FK_NKC * 64 + FK_WARPS * FK_NKC) + ((ntile + 15) & ~(size_t)15);
}

}  // namespace

extern "C" {

int mistra_fastkmt_device(int64_t ncell, const mistra_fastkmt_args *d_a, void *stream)
{
  int rc = check(ncell, d_a);
  if (rc) return rc;
  if (ncell == 0) return 0;
  int dev = -1, sms = 0;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  CKW(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = smem_bytes(d_a);
  if (smem > 220 * 1024) return mistra_internal_fail(MISTRA_KPP_EINVAL, "particle grid too large for shared memory");
  if (!g_attr[dev]) {
    CKW(cudaFuncSetAttribute(fastkmt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    g_attr[dev] = true;
  }
  long long blocks = sms;
  if (blocks > ncell) blocks = ncell;
  fastkmt_kernel<<<(int)blocks, FK_THREADS, smem, (cudaStream_t)stream>>>(ncell, *d_a);
  CKW(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mistra_fastkmt(int64_t ncell, const mistra_fastkmt_args *a, void *stream)
{
  int rc = check(ncell, a);
  if (rc) return rc;
  if (ncell == 0) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, nka = a->nka, nkt = a->nkt, nkc = a->nkc, ns = a->nspec;
  struct Item { const void *h; size_t bytes; bool in, out; void **slot; };
  mistra_fastkmt_args d = *a;
  std::vector<Item> items = {
      {a->lex, (size_t)a->nx * 4, true, false, (void **)&d.lex}, {a->kw, nka * 4, true, false, (void **)&d.kw},
      {a->rq, nka * nkt * 8, true, false, (void **)&d.rq}, {a->ff, n * nka * nkt * 8, true, false, (void **)&d.ff},
      {a->freep, n * 8, true, false, (void **)&d.freep}, {a->t, n * 8, true, false, (void **)&d.t},
      {a->p, n * 8, true, false, (void **)&d.p}, {a->cw, n * nkc * 8, true, false, (void **)&d.cw},
      {a->cm, n * nkc * 8, true, false, (void **)&d.cm}, {a->alpha, n * ns * 8, true, false, (void **)&d.alpha},
      {a->vmean, n * ns * 8, true, false, (void **)&d.vmean},
      {a->xkmt, n * nkc * ns * 8, true, true, (void **)&d.xkmt}, {a->vt, n * nkc * 8, true, true, (void **)&d.vt}};
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
  if ((rc = mistra_fastkmt_device(ncell, &d, stream))) return rc;
  for (auto &it : items)
    if (it.out) CKW(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKW(cudaStreamSynchronize(st));
  return 0;
}

int64_t mistra_fastkmt_launch_count(void) { return g_launches.load(); }

}  // extern "C"
