// SUBROUTINE difc on the device (include/mistra_difc.h): CUDA kernels + C-ABI entries.
// Role in the reference: /root/reference/src/str.f90:3271-3445 for every column of an ensemble.
//
// Two kernels.  difc_coef_kernel: one thread per column runs the matrix recurrence once
// (str.f90:3334-3343: xc, xd, xe and the subsidence Courant number c) into a scratch array that all
// species of the column share.  difc_solve_kernel: one thread per (column, species); a CTA covers
// 128 consecutive species of one column, so every access to the species arrays is a coalesced row
// segment and the coefficients are CTA-uniform (staged once in shared memory).  The forward
// elimination stores xf(k) in place of s(k) (the value is dead after it is read), the backward
// substitution reads it back - last written, first read, so the re-read is served by L1/L2 - and the
// subsidence step of level k is applied in the same sweep as soon as the diffused values of levels k
// and k+1 are known (str.f90:3354-3356 reads s(k+1) before it is updated), so each level is written
// once with its final value.  HBM-bound: 16 B per level and species.  No FMA contraction (build.py):
// bit-identical to the reference order.
#include "../../include/mistra_difc.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int DIFC_THREADS = 128;
constexpr int DIFC_MAXN = 512;

// coef [ncol][4][n]: xc, xd, xe, c (index 0 of each unused, as in the reference).  One warp per
// column: xa, xc, xb and c do not depend on the recurrence and are formed by all lanes (coalesced
// loads); lane 0 then runs the chain xd(k) = xb(k) - xc(k) xe(k-1), xe(k) = xa(k) / xd(k) out of shared
// memory and the warp writes the four rows back coalesced.
constexpr int COEF_WARPS = 4;
__global__ void __launch_bounds__(COEF_WARPS * 32) difc_coef_kernel(long long ncol, int n, double dt,
                                                                    const double *__restrict__ atkh,
                                                                    const double *__restrict__ w,
                                                                    const double *__restrict__ detw,
                                                                    const double *__restrict__ deta,
                                                                    double *__restrict__ coef)
{
  extern __shared__ double sm[];                           // per warp: xa, xb, xc, xd, xe [n]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long col = (long long)blockIdx.x * COEF_WARPS + warp;
  if (col >= ncol) return;
  double *xa = sm + (size_t)warp * 5 * n, *xb = xa + n, *xc = xb + n, *xd = xc + n, *xe = xd + n;
  const double *ak = atkh + col * n, *wk = w + col * n;
  double *o = coef + col * 4 * n;
  const int nm = n - 1;
  for (int k = lane; k < nm; k += 32) xa[k] = ak[k] * dt / (detw[k] * deta[k]);      // str.f90:3334, 3337
  __syncwarp();
  for (int k = lane; k <= nm; k += 32) {
    double xck = 0.0, ck = 0.0;
    if (k >= 1 && k < nm) {
      xck = xa[k - 1] * detw[k - 1] / detw[k];             // 3338
      xb[k] = 1.0 + xa[k] + xck;                           // 3339
      ck = wk[k] * dt / deta[k];                           // 3342
    }
    xc[k] = xck;
    o[3 * n + k] = ck;
  }
  __syncwarp();
  if (lane == 0) {
    double xe_prev = 0.0;                                  // xe(1) = 0
    xd[0] = 1.0; xe[0] = 0.0;
    for (int k = 1; k < nm; ++k) {
      const double xdk = xb[k] - xc[k] * xe_prev;          // 3340
      xe_prev = xa[k] / xdk;                               // 3341
      xd[k] = xdk; xe[k] = xe_prev;
    }
    xd[nm] = 1.0; xe[nm] = 0.0;
  }
  __syncwarp();
  for (int k = lane; k <= nm; k += 32) { o[k] = xc[k]; o[n + k] = xd[k]; o[2 * n + k] = xe[k]; }
}

// x / d, IEEE-exact for every input, without the slow path of the division for a zero numerator: when
// x == 0 and d is positive and finite (air density, the pivots xd >= 1) the quotient is x itself, sign
// of the zero included; those lanes divide 1 / d instead and the result is selected - no branch, so the
// empty bins of a sparse spectrum neither diverge nor drag their warp through the slow path (measured:
// 2.9 ms -> see DESIGN 5.8 for 2000 columns with a third of the species absent).
__device__ __forceinline__ double div_pos(double x, double d)
{
  // both tests on the integer pipe (FP64 compares share the half-rate FP64 pipe with the arithmetic):
  // x == +-0, and d a positive normal number (a denormal d simply takes the division)
  const bool z = (((__double2hiint(x) & 0x7fffffff) | __double2loint(x)) == 0) &
                 ((unsigned)(__double2hiint(d) - 0x00100000) < 0x7fe00000u);
  double xs = z ? 1.0 : x;
  asm("" : "+d"(xs));                  // opaque: keeps the compiler from dividing the original numerator and selecting afterwards
  const double q = xs / d;
  return z ? x : q;
}

struct DifcFields {                                        // the species of all arrays as one virtual row
  double *s[MISTRA_DIFC_MAXFIELDS];
  int row[MISTRA_DIFC_MAXFIELDS], first[MISTRA_DIFC_MAXFIELDS + 1];
  int nfield;
};

// DIFP = false: difc (species amounts s, mixing ratio s / am3).  DIFP = true: difp (str.f90:3137-3265):
// the particle spectrum is divided by rho at every level first (top level included, which therefore
// is rewritten as ff / rho * rho), the recurrences run on the ratio, the subsidence on the product.
template <bool DIFP, int DIFC_U>
__global__ void __launch_bounds__(DIFC_THREADS) difc_solve_kernel(int n, const double *__restrict__ coef,
                                                                  const double *__restrict__ am3, DifcFields fl,
                                                                  double *__restrict__ partial)
{
  extern __shared__ double sm[];                           // xc, xd, xe, c, am3 of the column: [5][n]; DIFP: + [4][n]
  const long long col = blockIdx.y;
  const double *cf = coef + col * 4 * n;
  for (int q = threadIdx.x; q < 4 * n; q += DIFC_THREADS) sm[q] = cf[q];
  for (int q = threadIdx.x; q < n; q += DIFC_THREADS) sm[4 * n + q] = am3[col * n + q];
  __syncthreads();
  const int j = blockIdx.x * DIFC_THREADS + threadIdx.x;
  const bool act = j < fl.first[fl.nfield];
  if (!DIFP && !act) return;                               // DIFP: idle lanes stay for the level sums (they add 0)
  int f = 0;
#pragma unroll
  for (int q = 1; q < MISTRA_DIFC_MAXFIELDS; ++q)
    if (q < fl.nfield && j >= fl.first[q]) f = q;
  const int row = fl.row[f];
  const double *xc = sm, *xd = sm + n, *xe = sm + 2 * n, *c = sm + 3 * n, *am = sm + 4 * n;
  double *s = fl.s[f] + (size_t)col * n * row + (act ? j - fl.first[f] : 0);
  const int nm = n - 1;
  // The sweeps run in batches of DIFC_U levels: the loads of the next batch are issued before the
  // dependent arithmetic and the stores of the current one (the stores may alias as far as the
  // compiler knows), so a thread waits for memory once per batch instead of once per level.
  double v[DIFC_U], vn[DIFC_U];
  double xf = div_pos(act ? s[(size_t)row] : 1.0, am[1]);  // xf(1) = s(j,2)/am3(2)
#pragma unroll
  for (int i = 0; i < DIFC_U; ++i) v[i] = (act && 1 + i < nm) ? s[(size_t)(1 + i) * row] : 1.0;
  for (int k = 1; k < nm; k += DIFC_U) {                   // forward elimination, str.f90:3348-3350 / 3230-3232
#pragma unroll
    for (int i = 0; i < DIFC_U; ++i) vn[i] = (act && k + DIFC_U + i < nm) ? s[(size_t)(k + DIFC_U + i) * row] : 1.0;
#pragma unroll
    for (int i = 0; i < DIFC_U; ++i)
      if (k + i < nm) {
        xf = div_pos(div_pos(v[i], am[k + i]) + xc[k + i] * xf, xd[k + i]);
        if (act) s[(size_t)(k + i) * row] = xf;
      }
#pragma unroll
    for (int i = 0; i < DIFC_U; ++i) v[i] = vn[i];
  }
  // backward: levels nm-1 .. 1; v[i] = xf(k - i)
#pragma unroll
  for (int i = 0; i < DIFC_U; ++i) v[i] = (act && nm - 1 - i >= 1) ? s[(size_t)(nm - 1 - i) * row] : 1.0;
  if (!DIFP) {
    double up = s[(size_t)nm * row];                       // s(j,n): boundary value, diffused value of level k+1 below
    for (int k = nm - 1; k >= 1; k -= DIFC_U) {            // back substitution 3351-3353 + subsidence 3354-3356
#pragma unroll
      for (int i = 0; i < DIFC_U; ++i) vn[i] = (k - DIFC_U - i >= 1) ? s[(size_t)(k - DIFC_U - i) * row] : 0.0;
#pragma unroll
      for (int i = 0; i < DIFC_U; ++i)
        if (k - i >= 1) {
          const double sd = (div_pos(xe[k - i] * up, am[k - i + 1]) + v[i]) * am[k - i];
          s[(size_t)(k - i) * row] = sd - c[k - i] * (up - sd);
          up = sd;
        }
#pragma unroll
      for (int i = 0; i < DIFC_U; ++i) v[i] = vn[i];
    }
  } else {
    // fsum(k), str.f90:3248-3255: the level sums of this CTA's grid points ride along - a fixed butterfly
    // over the lanes per level, the four warps in order below, the CTAs in order in difp_fsum_kernel
    double *s_part = sm + 5 * n;                           // [4 warps][n]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double upr = div_pos(act ? s[(size_t)nm * row] : 1.0, am[nm]);   // ff(n) / rho(n), 3210-3212
    double upm = upr * am[nm];                             // ... * rho(n), 3238-3240
    if (act) s[(size_t)nm * row] = upm;
    {
      double w = act ? upm : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) w = w + __shfl_xor_sync(0xffffffffu, w, off);
      if (lane == 0) s_part[warp * n + nm] = w;
    }
    for (int k = nm - 1; k >= 1; k -= DIFC_U) {            // 3234-3236, 3238-3240, 3243-3246
#pragma unroll
      for (int i = 0; i < DIFC_U; ++i) vn[i] = (act && k - DIFC_U - i >= 1) ? s[(size_t)(k - DIFC_U - i) * row] : 1.0;
#pragma unroll
      for (int i = 0; i < DIFC_U; ++i)
        if (k - i >= 1) {
          const double xr = xe[k - i] * upr + v[i];
          const double sp = xr * am[k - i];
          const double fin = sp - c[k - i] * (upm - sp);
          if (act) s[(size_t)(k - i) * row] = fin;
          upr = xr; upm = sp;
          double w = act ? fin : 0.0;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) w = w + __shfl_xor_sync(0xffffffffu, w, off);
          if (lane == 0) s_part[warp * n + k - i] = w;
        }
#pragma unroll
      for (int i = 0; i < DIFC_U; ++i) v[i] = vn[i];
    }
    __syncthreads();
    double *po = partial + ((size_t)col * gridDim.x + blockIdx.x) * n;
    for (int k = 1 + threadIdx.x; k <= nm; k += DIFC_THREADS) {
      double t = s_part[k];
#pragma unroll
      for (int q = 1; q < DIFC_THREADS / 32; ++q) t = t + s_part[q * n + k];
      po[k] = t;
    }
  }
}

// fsum(k) = sum over the CTAs of a column of their level sums, in CTA order (str.f90:3248-3255; the
// reference keeps one running sum over the nka * nkt points).
__global__ void __launch_bounds__(128) difp_fsum_kernel(int n, int nblk, const double *__restrict__ partial, double *fsum)
{
  const long long col = blockIdx.x;
  for (int k = 1 + threadIdx.x; k < n; k += 128) {
    const double *p = partial + (size_t)col * nblk * n + k;
    double t = p[0];
    for (int b = 1; b < nblk; ++b) t = t + p[(size_t)b * n];
    fsum[col * n + k] = t;
  }
}

std::mutex g_mu;
std::atomic<long long> g_launches{0};
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_scratch[16], g_coef[16];

#define CKW(call)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return mistra_internal_fail(e_ == cudaErrorMemoryAllocation ? MISTRA_KPP_ENOMEM   \
                                  : (e_ == cudaErrorNoDevice ? MISTRA_KPP_ENODEVICE     \
                                                             : MISTRA_KPP_ECUDA),       \
                                  std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

// dynamic shared memory of the solve kernel: the column's coefficients, padded so that at most
// `resident` CTAs (MISTRA_DIFC_CTAS_PER_SM, default 16) share an SM - fewer resident species keep the
// xf values that wait for the backward sweep inside the L2.
size_t solve_smem(int n)
{
  static const int resident = [] {
    const char *e = getenv("MISTRA_DIFC_CTAS_PER_SM");
    const int v = e ? atoi(e) : 16;
    return v < 1 ? 1 : (v > 16 ? 16 : v);
  }();
  const size_t need = 9 * (size_t)n * sizeof(double), cap = (size_t)(224 * 1024) / resident - 1024;
  return need > cap ? need : cap;
}

constexpr int DIFC_BATCH = 8;    // levels per batch of the sweeps (16 measured slower: 96 registers, 5 CTAs per SM)

bool g_attr[16] = {};
int set_attrs(int dev)
{
  if (g_attr[dev]) return 0;
  CKW(cudaFuncSetAttribute(difc_solve_kernel<false, DIFC_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  CKW(cudaFuncSetAttribute(difc_solve_kernel<true, DIFC_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  CKW(cudaFuncSetAttribute(difc_coef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  g_attr[dev] = true;
  return 0;
}

int check(int64_t ncol, const mistra_difc_args *a)
{
  if (ncol < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncol < 0");
  if (!a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null arguments");
  if (a->n < 3 || a->n > DIFC_MAXN) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad level count (3 <= n <= 512)");
  if (a->nfield < 0 || a->nfield > MISTRA_DIFC_MAXFIELDS)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad field count");
  if (!a->detw || !a->deta) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null grid array");
  if (ncol > 0 && (!a->atkh || !a->w || !a->am3)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  for (int f = 0; f < a->nfield; ++f) {
    const mistra_difc_field &fl = a->field[f];
    if (fl.row < 1 || fl.nproc < 0 || fl.nproc > fl.row || (ncol > 0 && !fl.s))
      return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad field (row >= 1, 0 <= nproc <= row, non-null)");
  }
  return 0;
}

}  // namespace

extern "C" {

int mistra_difc_device(int64_t ncol, const mistra_difc_args *d_a, void *stream)
{
  int rc = check(ncol, d_a);
  if (rc) return rc;
  if (ncol == 0) return 0;
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = d_a->n;
  if (int rca = set_attrs(dev)) return rca;
  const size_t need = (size_t)ncol * 4 * n * sizeof(double);
  Scratch &cf = g_coef[dev];
  if (cf.bytes < need) {
    if (cf.p) { CKW(cudaDeviceSynchronize()); cudaFree(cf.p); cf.p = nullptr; cf.bytes = 0; }
    CKW(cudaMalloc(&cf.p, need));
    cf.bytes = need;
  }
  double *coef = (double *)cf.p;
  difc_coef_kernel<<<(unsigned)((ncol + COEF_WARPS - 1) / COEF_WARPS), COEF_WARPS * 32, COEF_WARPS * 5 * n * sizeof(double),
                     st>>>(ncol, n, d_a->dt, d_a->atkh, d_a->w, d_a->detw, d_a->deta, coef);
  CKW(cudaGetLastError());
  g_launches.fetch_add(1);
  DifcFields fl;
  fl.nfield = 0;
  fl.first[0] = 0;
  for (int f = 0; f < d_a->nfield; ++f) {
    if (d_a->field[f].nproc == 0) continue;
    fl.s[fl.nfield] = d_a->field[f].s;
    fl.row[fl.nfield] = d_a->field[f].row;
    fl.first[fl.nfield + 1] = fl.first[fl.nfield] + d_a->field[f].nproc;
    ++fl.nfield;
  }
  if (fl.nfield == 0) return 0;
  for (int64_t c0 = 0; c0 < ncol; c0 += 65535) {           // grid.y limit
    const int64_t nc = ncol - c0 < 65535 ? ncol - c0 : 65535;
    DifcFields fc = fl;
    for (int f = 0; f < fl.nfield; ++f) fc.s[f] = fl.s[f] + (size_t)c0 * n * fl.row[f];
    dim3 grid((fl.first[fl.nfield] + DIFC_THREADS - 1) / DIFC_THREADS, (unsigned)nc);
    difc_solve_kernel<false, DIFC_BATCH><<<grid, DIFC_THREADS, solve_smem(n), st>>>(n, coef + c0 * 4 * n, d_a->am3 + c0 * n,
                                                                                   fc, nullptr);
    CKW(cudaGetLastError());
    g_launches.fetch_add(1);
  }
  return 0;
}

int mistra_difc(int64_t ncol, const mistra_difc_args *a, void *stream)
{
  int rc = check(ncol, a);
  if (rc) return rc;
  if (ncol == 0) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nc = (size_t)ncol, n = a->n;
  struct Item { const void *h; size_t bytes; bool in, out; void **slot; };
  mistra_difc_args d = *a;
  std::vector<Item> items = {
      {a->atkh, nc * n * 8, true, false, (void **)&d.atkh}, {a->w, nc * n * 8, true, false, (void **)&d.w},
      {a->am3, nc * n * 8, true, false, (void **)&d.am3}, {a->detw, n * 8, true, false, (void **)&d.detw},
      {a->deta, n * 8, true, false, (void **)&d.deta}};
  for (int f = 0; f < a->nfield; ++f)
    items.push_back({a->field[f].s, nc * n * (size_t)a->field[f].row * 8, true, true, (void **)&d.field[f].s});
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
  if ((rc = mistra_difc_device(ncol, &d, stream))) return rc;
  for (auto &it : items)
    if (it.out) CKW(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKW(cudaStreamSynchronize(st));
  return 0;
}

int mistra_difp_device(int64_t ncol, const mistra_difp_args *d_a, void *stream)
{
  if (ncol < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncol < 0");
  if (!d_a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null arguments");
  if (d_a->n < 3 || d_a->n > DIFC_MAXN || d_a->row < 1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (3 <= n <= 512, row >= 1)");
  if (!d_a->detw || !d_a->deta) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null grid array");
  if (ncol > 0 && (!d_a->atkh || !d_a->w || !d_a->rho || !d_a->ff || !d_a->fsum))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = d_a->n;
  if (int rca = set_attrs(dev)) return rca;
  const size_t nblk = ((size_t)d_a->row + DIFC_THREADS - 1) / DIFC_THREADS;
  const size_t ncmax = ncol < 65535 ? (size_t)ncol : 65535;
  const size_t need = ((size_t)ncol * 4 * n + ncmax * nblk * n) * sizeof(double);   // coefficients + level sums per CTA
  Scratch &cf = g_coef[dev];
  if (cf.bytes < need) {
    if (cf.p) { CKW(cudaDeviceSynchronize()); cudaFree(cf.p); cf.p = nullptr; cf.bytes = 0; }
    CKW(cudaMalloc(&cf.p, need));
    cf.bytes = need;
  }
  double *coef = (double *)cf.p, *partial = coef + (size_t)ncol * 4 * n;
  difc_coef_kernel<<<(unsigned)((ncol + COEF_WARPS - 1) / COEF_WARPS), COEF_WARPS * 32, COEF_WARPS * 5 * n * sizeof(double),
                     st>>>(ncol, n, d_a->dt, d_a->atkh, d_a->w, d_a->detw, d_a->deta, coef);
  CKW(cudaGetLastError());
  g_launches.fetch_add(1);
  for (int64_t c0 = 0; c0 < ncol; c0 += 65535) {           // grid.y limit
    const int64_t nc = ncol - c0 < 65535 ? ncol - c0 : 65535;
    dim3 grid((unsigned)nblk, (unsigned)nc);
    double *ffc = d_a->ff + (size_t)c0 * n * d_a->row;
    DifcFields fc;
    fc.nfield = 1; fc.s[0] = ffc; fc.row[0] = d_a->row; fc.first[0] = 0; fc.first[1] = d_a->row;
    difc_solve_kernel<true, DIFC_BATCH><<<grid, DIFC_THREADS, solve_smem(n), st>>>(n, coef + c0 * 4 * n, d_a->rho + c0 * n,
                                                                                  fc, partial);
    CKW(cudaGetLastError());
    difp_fsum_kernel<<<(unsigned)nc, 128, 0, st>>>(n, (int)grid.x, partial, d_a->fsum + c0 * n);
    CKW(cudaGetLastError());
    g_launches.fetch_add(2);
  }
  return 0;
}

int mistra_difp(int64_t ncol, const mistra_difp_args *a, void *stream)
{
  if (ncol < 0 || !a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  if (a->n < 3 || a->n > DIFC_MAXN || a->row < 1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (3 <= n <= 512, row >= 1)");
  if (!a->detw || !a->deta || (ncol > 0 && (!a->atkh || !a->w || !a->rho || !a->ff || !a->fsum)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nc = (size_t)ncol, n = a->n;
  struct Item { const void *h; size_t bytes; bool in, out; void **slot; };
  mistra_difp_args d = *a;
  std::vector<Item> items = {
      {a->atkh, nc * n * 8, true, false, (void **)&d.atkh}, {a->w, nc * n * 8, true, false, (void **)&d.w},
      {a->rho, nc * n * 8, true, false, (void **)&d.rho}, {a->detw, n * 8, true, false, (void **)&d.detw},
      {a->deta, n * 8, true, false, (void **)&d.deta},
      {a->ff, nc * n * (size_t)a->row * 8, true, true, (void **)&d.ff}, {a->fsum, nc * n * 8, true, true, (void **)&d.fsum}};
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
  int rc = mistra_difp_device(ncol, &d, stream);
  if (rc) return rc;
  for (auto &it : items)
    if (it.out) CKW(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKW(cudaStreamSynchronize(st));
  return 0;
}

int64_t mistra_difc_launch_count(void) { return g_launches.load(); }

}  // extern "C"
