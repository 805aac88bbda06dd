// SUBROUTINE fast_k_mt_a / fast_k_mt_t on the device (include/mistra_fastkmt.h): CUDA kernel + C-ABI
// entries.  Role in the reference: the layer loop of /root/reference/src/kpp.f90:2820-2945 (and
// 2558-2674), with FUNCTION vterm (str.f90:2793-2864).
//
// Mapping: one persistent CTA (512 threads) per SM, a layer at a time, the next layer's 70 x 70
// spectrum arriving by 16-byte asynchronous copies into the second half of a double buffer while the
// current one is integrated.  The grid tables - radius rqm and the chemistry bin of every grid point
// (aerosol / droplet part by kw, small / large classes by ka, 255 = not summed) - are put in shared
// memory once.  Per layer: (1) the grid points that contribute are compacted, in grid order, into two
// index lists (ballot + prefix over the warps): points with ff != 0 for the sedimentation sums, and
// those of them whose bin has chemistry (cm > 0) for the transfer coefficients - every other point
// adds exactly +0 in the reference; (2) a thread per listed point: q = rqm / freep and the vterm
// term; (3) lanes = exchanged species (l and l + 32), warps stride over the listed points, three
// points per iteration so that every thread has six independent FP64 divisions in flight; each
// thread keeps its species' sums of the four bins in registers.  Partial sums are combined over the
// warps in warp order.  Real spectra populate a few per cent of the grid, then the kernel streams
// ff (39.2 kB per layer); with every point populated it is FP64-bound: (nx + 1) divisions per point.
// No FMA contraction (build.py).
#include "tma_bulk.h"
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

struct FkSmem {
  double *f0, *r, *q, *red, *vt, *sc;      // f0: [2][npad]
  int npad;
  unsigned short *l1, *l2;
  unsigned char *kc;
  int *cnt;
};

__device__ __forceinline__ FkSmem fk_carve(double *smem, int ntile)
{
  const int npad = (ntile + 1) & ~1, lpad = (ntile + 7) & ~7;
  FkSmem s;
  s.npad = npad;
  s.f0 = smem;                              // [2][npad]: the layer's spectrum, double-buffered
  s.r = s.f0 + 2 * npad;                     // [ntile] rqm = rq * 1e-6 (kpp.f90:2806)
  s.q = s.r + npad;                         // [n2] rqm / freep(k) of the points of list 2
  s.red = s.q + npad;                       // [FK_WARPS][FK_NKC][64] partial sums of the transfer coefficients
  s.vt = s.red + FK_WARPS * FK_NKC * 64;    // [FK_WARPS][FK_NKC]
  s.sc = s.vt + FK_WARPS * FK_NKC;          // [2][16] layer scalars: cm[4], cw[4], freep, t, p (double-buffered)
  s.l1 = (unsigned short *)(s.sc + 32);     // points with ff != 0 (sedimentation sums)
  s.l2 = s.l1 + lpad;                       // ... whose bin has chemistry (transfer coefficients)
  s.kc = (unsigned char *)(s.l2 + lpad);    // [ntile] bin of the point, 255 = none
  s.cnt = (int *)(s.kc + lpad);             // [FK_WARPS] list lengths per warp (two 16-bit counts)
  return s;
}

// returns true if the tile travels by a bulk (TMA) copy whose completion is signalled on `bar`
__device__ __forceinline__ bool fk_fetch(double *dst, double *sc, const mistra_fastkmt_args &a, long long c, int ntile,
                                         unsigned long long *bar)
{
  const double *gf = a.ff + (size_t)c * ntile;
  const bool bulk = tma::bulk_ok(gf, (size_t)ntile * 8);
  if (bulk) {
    // the layer's spectrum is one contiguous 39.2 kB block: ONE bulk copy instead of 2450 16-byte copies
    if (threadIdx.x == 0) tma::bulk_load(dst, gf, (unsigned)(ntile * 8), bar);
  } else if ((ntile & 1) == 0 && ((reinterpret_cast<unsigned long long>(gf) & 15ull) == 0ull)) {
    for (int q = threadIdx.x; q < (ntile >> 1); q += FK_THREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst + 2 * q)),
                   "l"(gf + 2 * q) : "memory");
  } else {
    for (int q = threadIdx.x; q < ntile; q += FK_THREADS)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst + q)),
                   "l"(gf + q) : "memory");
  }
  // the layer's scalars travel with the tile: cm[0..3], cw[0..3], freep, t, p
  const int tq = threadIdx.x;
  const double *src = nullptr;
  if (tq < 4) { if (tq < a.nkc) src = a.cm + c * a.nkc + tq; }
  else if (tq < 8) { if (tq - 4 < a.nkc) src = a.cw + c * a.nkc + (tq - 4); }
  else if (tq == 8) src = a.freep + c;
  else if (tq == 9) src = a.t + c;
  else if (tq == 10) src = a.p + c;
  if (src)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(sc + tq)), "l"(src)
                 : "memory");
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  return bulk;
}

__global__ void __launch_bounds__(FK_THREADS, 1) fastkmt_kernel(long long ncell, mistra_fastkmt_args a)
{
  extern __shared__ __align__(16) double smem[];
  const int nka = a.nka, nkt = a.nkt, ntile = nka * nkt;
  const FkSmem s = fk_carve(smem, ntile);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double z4pi3 = 4.0 * 3.1415926535897932 / 3.0;   // kpp.f90:2758
  const int nx = a.nx, nspec = a.nspec, nkc = a.nkc;

  __shared__ unsigned long long s_bar[2];    // completion barriers of the two tile buffers' bulk copies
  if (threadIdx.x == 0) { tma::mbar_init(&s_bar[0], 1); tma::mbar_init(&s_bar[1], 1); }
  __syncthreads();
  unsigned phase = 0;                        // bit b: parity of buffer b's next completion
  bool bulk_cur = false, bulk_nxt = false;
  if ((long long)blockIdx.x < ncell) bulk_cur = fk_fetch(s.f0, s.sc, a, blockIdx.x, ntile, &s_bar[0]);
  for (int q = threadIdx.x; q < ntile; q += FK_THREADS) {
    const int ia = q / nkt, jt = q - ia * nkt;
    s.r[q] = a.rq[q] * 1.e-6;
    int kc = (jt < a.kw[ia] ? 0 : 2) + (ia < a.ka ? 0 : 1);          // kpp.f90:2869-2901
    if (kc >= a.nkc_l || (ia < a.ial - 1 && ia < a.ka)) kc = 255;    // ifeed == 2 drops the first small class
    s.kc[q] = (unsigned char)kc;
  }
  const int l0 = lane, l1 = lane + 32;
  const int sp0 = l0 < nx ? a.lex[l0] - 1 : -1, sp1 = l1 < nx ? a.lex[l1] - 1 : -1;

  int buf = 0;
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x, buf ^= 1) {
    __syncthreads();                         // the other buffer, the lists and the partial sums are free again
    const long long cn = c + gridDim.x;
    if (cn < ncell) bulk_nxt = fk_fetch(s.f0 + (buf ^ 1) * s.npad, s.sc + (buf ^ 1) * 16, a, cn, ntile, &s_bar[buf ^ 1]);
    // the species' accommodation coefficient and mean speed while the tile is in flight
    // (loads issued here, consumed in step 3).  Idle lanes: 1 / (q + 1), never stored - a zero numerator
    // would send the whole warp down the slow path of the division.
    double vm0 = 1.0, vm1 = 1.0, al0 = 0.75, al1 = 0.75;
    if (sp0 >= 0) { al0 = a.alpha[c * nspec + sp0]; vm0 = a.vmean[c * nspec + sp0]; }
    if (sp1 >= 0) { al1 = a.alpha[c * nspec + sp1]; vm1 = a.vmean[c * nspec + sp1]; }
    if (cn < ncell) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (bulk_cur) {
      tma::mbar_wait(&s_bar[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
    }
    bulk_cur = bulk_nxt;
    __syncthreads();
    const double *sf = s.f0 + buf * s.npad, *sc = s.sc + buf * 16;
    const double freep = sc[8], tk = sc[9], pk = sc[10];
    unsigned onmask = 0;                     // bins with chemistry (cm > 0, kpp.f90:2826)
    for (int kc = 0; kc < a.nkc_l; ++kc)
      if (sc[kc] > 0.0) onmask |= 1u << kc;

    // (1) compaction in grid order: warp w owns the contiguous segment [w * seg, (w + 1) * seg) of the grid;
    // first pass counts, one exclusive scan over the warps, second pass writes the indices
    const int seg = ((ntile + FK_WARPS - 1) / FK_WARPS + 31) & ~31;
    const int q0 = warp * seg;
    const unsigned below = (1u << lane) - 1u;
    int c1 = 0, c2 = 0;
    for (int base = q0; base < q0 + seg; base += 32) {
      const int q = base + lane;
      bool act1 = false, act2 = false;
      if (q < ntile) {
        const int kc = s.kc[q];
        act1 = kc != 255 && sf[q] != 0.0;
        act2 = act1 && ((onmask >> kc) & 1u);
      }
      c1 += __popc(__ballot_sync(0xffffffffu, act1));
      c2 += __popc(__ballot_sync(0xffffffffu, act2));
    }
    if (lane == 0) s.cnt[warp] = c1 | (c2 << 16);
    __syncthreads();
    int n1, n2, o1, o2;
    {
      const int mine = lane < FK_WARPS ? s.cnt[lane] : 0;   // both counts < 2^16
      int inc = mine;
#pragma unroll
      for (int off = 1; off < FK_WARPS; off <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += up;
      }
      const int tot = __shfl_sync(0xffffffffu, inc, FK_WARPS - 1);
      const int exc = __shfl_sync(0xffffffffu, inc - mine, warp);
      n1 = tot & 0xffff; n2 = tot >> 16;
      o1 = exc & 0xffff; o2 = exc >> 16;
    }
    for (int base = q0; base < q0 + seg; base += 32) {
      const int q = base + lane;
      bool act1 = false, act2 = false;
      if (q < ntile) {
        const int kc = s.kc[q];
        act1 = kc != 255 && sf[q] != 0.0;
        act2 = act1 && ((onmask >> kc) & 1u);
      }
      const unsigned b1 = __ballot_sync(0xffffffffu, act1), b2 = __ballot_sync(0xffffffffu, act2);
      if (act1) s.l1[o1 + __popc(b1 & below)] = (unsigned short)q;
      if (act2) s.l2[o2 + __popc(b2 & below)] = (unsigned short)q;
      o1 += __popc(b1); o2 += __popc(b2);
    }
    __syncthreads();

    // (2) q = rqm / freep of list 2; sedimentation sums over list 1 (kpp.f90:2924-2927)
    for (int i = threadIdx.x; i < n2; i += FK_THREADS) s.q[i] = s.r[s.l2[i]] / freep;
    {
      double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
      for (int i = threadIdx.x; i < n1; i += FK_THREADS) {
        const int q = s.l1[i];
        const double rqq = s.r[q];
        const double xvs = vterm(rqq, tk, pk);
        const double term = rqq * rqq * rqq * xvs * sf[q] * 1.e6;
        const int kc = s.kc[q];
        if (kc == 0) v0 = v0 + term; else if (kc == 1) v1 = v1 + term; else if (kc == 2) v2 = v2 + term; else v3 = v3 + term;
      }
      v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
      if (lane == 0) { s.vt[warp * 4 + 0] = v0; s.vt[warp * 4 + 1] = v1; s.vt[warp * 4 + 2] = v2; s.vt[warp * 4 + 3] = v3; }
    }
    __syncthreads();

    // (3) transfer coefficients (kpp.f90:2913-2922); warp-uniform control flow
    if (n2 > 0) {
      const double x10 = (al0 > 0.0) ? 4. / (3. * al0) : 0.0;   // kpp.f90:2890
      const double x11 = (al1 > 0.0) ? 4. / (3. * al1) : 0.0;
      double a00 = 0.0, a01 = 0.0, a02 = 0.0, a03 = 0.0, a10 = 0.0, a11 = 0.0, a12 = 0.0, a13 = 0.0;
      for (int i = warp; i < n2; i += 3 * FK_WARPS) {
        const int ib = i + FK_WARPS, ic = i + 2 * FK_WARPS;
        const bool vb = ib < n2, vc = ic < n2;
        const int pA = s.l2[i], pB = vb ? s.l2[ib] : pA, pC = vc ? s.l2[ic] : pA;
        const double fA = sf[pA], fB = vb ? sf[pB] : 0.0, fC = vc ? sf[pC] : 0.0;    // a missing point adds +0
        const double rA = s.r[pA], rB = s.r[pB], rC = s.r[pC];
        const double qA = s.q[i], qB = s.q[vb ? ib : i], qC = s.q[vc ? ic : i];
        const int kcA = s.kc[pA], kcB = s.kc[pB], kcC = s.kc[pC];
        const double xA0 = vm0 / (qA + x10), xA1 = vm1 / (qA + x11);
        const double xB0 = vm0 / (qB + x10), xB1 = vm1 / (qB + x11);
        const double xC0 = vm0 / (qC + x10), xC1 = vm1 / (qC + x11);
        const double tA0 = xA0 * rA * rA * fA * 1.e6, tA1 = xA1 * rA * rA * fA * 1.e6;
        const double tB0 = xB0 * rB * rB * fB * 1.e6, tB1 = xB1 * rB * rB * fB * 1.e6;
        const double tC0 = xC0 * rC * rC * fC * 1.e6, tC1 = xC1 * rC * rC * fC * 1.e6;
        FK_ACC(kcA, tA0, tA1);
        FK_ACC(kcB, tB0, tB1);
        FK_ACC(kcC, tC0, tC1);
      }
      double *r = s.red + warp * (FK_NKC * 64);
      r[0 * 64 + l0] = a00; r[1 * 64 + l0] = a01; r[2 * 64 + l0] = a02; r[3 * 64 + l0] = a03;
      r[0 * 64 + l1] = a10; r[1 * 64 + l1] = a11; r[2 * 64 + l1] = a12; r[3 * 64 + l1] = a13;
    }
    __syncthreads();

    // results (kpp.f90:2931-2937): warps in order
    if (threadIdx.x < FK_NKC * 64) {
      const int kc = threadIdx.x >> 6, l = threadIdx.x & 63;
      if (kc < a.nkc_l && l < nx && ((onmask >> kc) & 1u)) {
        const double cw = sc[4 + kc];
        if (cw > 0.0) {
          double sum = 0.0;
          if (n2 > 0)
            for (int w = 0; w < FK_WARPS; ++w) sum = sum + s.red[w * (FK_NKC * 64) + kc * 64 + l];
          a.xkmt[((size_t)c * nkc + kc) * nspec + a.lex[l] - 1] = z4pi3 / cw * sum;
        }
      }
    } else if (threadIdx.x < FK_NKC * 64 + FK_NKC) {
      const int kc = threadIdx.x - FK_NKC * 64;
      if (kc < a.nkc_l) {
        const double cw = sc[4 + kc];
        if (cw > 0.0) {
          double sum = 0.0;
          for (int w = 0; w < FK_WARPS; ++w) sum = sum + s.vt[w * 4 + kc];
          a.vt[c * nkc + kc] = z4pi3 / cw * sum;
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
  const size_t ntile = (size_t)a->nka * a->nkt, npad = (ntile + 1) & ~(size_t)1, lpad = (ntile + 7) & ~(size_t)7;
  return sizeof(double) * (4 * npad + FK_WARPS * FK_NKC * 64 + FK_WARPS * FK_NKC + 32) + 2 * 2 * lpad + lpad +
         sizeof(int) * 2 * FK_WARPS;
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
  if (smem > 226 * 1024 || (size_t)d_a->nka * d_a->nkt > 65535)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "particle grid too large for shared memory (nka * nkt <= ~5300)");
  if (!g_attr[dev]) {
    CKW(cudaFuncSetAttribute(fastkmt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
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
