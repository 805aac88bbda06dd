// SUBROUTINE sedp / sedl / sedc on the device (include/mistra_sed.h): CUDA kernels + C-ABI entries.  Role in the
// reference: /root/reference/src/str.f90:2257-2411 (sedp), 2627-2787 (sedl), 2567-2596 (sedc, species loop), with
// advsed0 / advsed1 (5522-5691) and vterm (2793-2864).
//
// Mapping.  Every settling profile - one class (jt, ia) of the spectrum in sedp, one species (l, kc) of the aqueous
// arrays in sedl - is an independent 1-D advection problem over the levels 1..nf whose flux limiter is a recurrence
// from the top level down (fm(i-1) needs fm(i), str.f90:5655-5672), repeated for a data-dependent number of
// sub-steps.  One thread owns one profile; consecutive threads are consecutive species (sedl: one coalesced row per
// warp and level) or neighbouring settling classes (sedp).  A pass works IN PLACE on the profile in global memory:
// the five old values the polynomial fit needs slide through registers, the values entering that window are loaded
// four levels at a time and one chunk ahead of their use, and the new value of level i is stored as soon as fm(i-1)
// is known - neither psi nor the fit coefficients a0..a4 nor the fluxes are kept anywhere else, so the number of
// resident warps is bounded by registers only (the first version kept psi in shared memory: 800 B per thread,
// 8 warps per SM).  The first pass reads ff * detw, the last one writes psi / detw, passes in between leave psi itself
// in the array (profiles with one pass - almost all - touch each level once).  Courant numbers are formed on the way:
// in sedp from vterm at that level (the reference calls vterm again in every sub-step, str.f90:2364, and so does this
// kernel - rho_a and eta of the level are tabulated in shared memory), in sedl from the block's table cc(k) of the
// bin (all threads of a block share the bin, hence the number of sub-steps).  sedp first finds the classes that hold
// particles (the others are skipped by the reference, and are the bulk of a real spectrum), lists them by size and
// hands them to persistent blocks 64 at a time (see "sedp" below).  Its diagnostics need the classes in the
// reference's order (running sums, and x0 carried from one class to the next): the work kernel leaves x0 per class,
// sedp_diag_kernel (one block per column) forms x2 in parallel and adds up in order, one thread per sum.
// Roofline: HBM-side, ff is read once by the scan and the settling classes once more in and out; but a level of a
// settling profile costs ~200 FP64-pipe instructions per sub-step in a dependent chain (four IEEE divisions, five exact
// divisions by constants, vterm), so the work kernels are latency / FP64-bound.  No FMA contraction (build.py).
#include "../../include/mistra_sed.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int SEDL_T = 64;       // species per block
#ifndef SEDP_MINB
#define SEDP_MINB 10              // resident blocks per SM the compiler plans for (register budget)
#endif
#ifndef SEDL_MINB
#define SEDL_MINB 16              // measured (gpurun_out/r02_sed_minb_sweep.txt): more resident warps beat fewer spills
#endif
constexpr int SED_MAXN = 512;

__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }   // gfortran MIN / MAX
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }

// x / d without the slow path CUDA's division takes for a zero numerator (see difc_kernels.cu): exact for d a
// positive normal number
__device__ __forceinline__ double div_pos(double x, double d)
{
  const bool z = (((__double2hiint(x) & 0x7fffffff) | __double2loint(x)) == 0) &
                 ((unsigned)(__double2hiint(d) - 0x00100000) < 0x7fe00000u);
  double xs = z ? 1.0 : x;
  asm("" : "+d"(xs));
  const double q = xs / d;
  return z ? x : q;
}

// x / (D * 2^K) for D = 3 or 15, correctly rounded, without a division: with r = RN(1 / D) * 2^-K (the scaling is exact)
// q0 = RN(x * r) is within one ulp of the quotient, e = x - q0 * d is exact (one FMA), and RN(q0 + e * r) is the
// correctly rounded quotient (Markstein's theorem: r is the correctly rounded reciprocal of a divisor whose significand
// is not all ones, q0 a faithful quotient).  The theorem needs normal numbers throughout, hence the exponent window
// 2^-895 <= |x| < 2^897; everything else (zero, denormals, huge values, Inf / NaN) takes the zero shortcut or the IEEE
// division.  advsed1 divides by 24, 48, 1920, 384, 768 and 3840 at every level (str.f90:5613-5632): five of the six
// divisions of a level.  mistra_sed_divc_selftest compares it with the division on the device.
template <int D, int K>
__device__ __forceinline__ double div_const(double x)
{
  static_assert(D == 3 || D == 15, "divisor");
  const double d = (double)D * (double)(1 << K);
  const double r = __longlong_as_double(D == 3 ? 0x3FD5555555555555LL : 0x3FB1111111111111LL) * (1.0 / (double)(1 << K));
  const int hi = __double2hiint(x);
  const unsigned ex = ((unsigned)hi >> 20) & 0x7ffu;
  if (ex - 128u < 1792u) {
    const double q0 = __dmul_rn(x, r);
    const double e = __fma_rn(-q0, d, x);
    return __fma_rn(e, r, q0);
  }
  if (((hi & 0x7fffffff) | __double2loint(x)) == 0) return x;
  return x / d;
}

// FUNCTION vterm(a,t,p), str.f90:2793-2864, with rho_a = p/(r0*t) and eta = 3.7957d-06+4.9d-08*t of the level given
__device__ __forceinline__ double vterm_lev(double a, double t, double p, double rho_a, double eta)
{
  const double g = 9.80665, rhow = 1000.0;   // constants.f90
  const double b0 = -.318657e+1, b1 = .992696e+0, b2 = -.153193e-2, b3 = -.987059e-3, b4 = -.578878e-3,
               b5 = +.855176e-4, b6 = -.327815e-5;
  const double c1 = 2.0 * g / 9.0, c2 = 1.26, P0 = 101325, T0 = 293.15, lambda0 = 6.6e-8;
  const double c3 = c2 * lambda0 * P0 / T0, c4 = 32.0 * g / 3.0;
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

__device__ __forceinline__ double rho_air(double t, double p)
{
  const double gas_const = 8.3144743, M_air = 28.96546e-3;
  const double r0 = gas_const / M_air;
  return p / (r0 * t);
}
__device__ __forceinline__ double eta_air(double t) { return 3.7957e-06 + 4.9e-08 * t; }

// SUBROUTINE advsed1 (str.f90:5585-5691), one pass over a profile that lives in global memory.  0-based level i:
// ld(i), 1 <= i <= nf-2, is psi of the level before the pass, st(i, v) takes its new value; psi(1) = x1 (the caller's
// copy of psi(2), str.f90:2377) and psi(nf) = ytop are registers: ytop is updated, the new psi(1) is returned.
// cneg(i) = -c(i+1) of the reference, evaluated once per level from the top down.  The five old values of the
// polynomial fit slide through registers; the values entering the window are loaded four levels at a time, one chunk
// ahead of their use, so a thread waits for memory once per four levels at most.
template <class LD, class ST, class CNEG>
__device__ __forceinline__ double advsed1_pass(int nf, double x1, LD ld, ST st, CNEG cneg, double &ytop)
{
  auto yv = [&](int j) -> double { return j >= 1 ? ld(j) : (j == 0 ? x1 : 0.0); };
  double cl = cneg(nf - 2);
  const double yt = ytop, ytm = ld(nf - 2);
  double fm_up = dmin(yt, cl * (yt - (1.0 - cl) * (yt - ytm) * 0.5));          // fm(nf-1)
  ytop = yt - fm_up;
  double clm = cl;
  double yp2 = 0.0, yp1 = yt, y0 = ytm, ym1 = yv(nf - 3), ym2 = yv(nf - 4);
  auto step = [&](int i, double next) {
    cl = clm;
    clm = cneg(i - 1);
    double a0, a1, a2, a3, a4;
    if (i == 1 || i == nf - 2) {                                              // 5613-5617, 5628-5632
      a0 = div_const<3, 3>(26.0 * y0 - yp1 - ym1);
      a1 = (yp1 - ym1) / 16.0;
      a2 = div_const<3, 4>(yp1 + ym1 - 2.0 * y0);
      a3 = 0.0;
      a4 = 0.0;
    } else {                                                                  // 5619-5626
      a0 = div_const<15, 7>(9.0 * (yp2 + ym2) - 116.0 * (yp1 + ym1) + 2134.0 * y0);
      a1 = div_const<3, 7>(-5.0 * (yp2 - ym2) + 34.0 * (yp1 - ym1));
      a2 = div_const<3, 7>(-yp2 + 12.0 * (yp1 + ym1) - 22.0 * y0 - ym2);
      a3 = div_const<3, 8>(yp2 - 2.0 * (yp1 - ym1) - ym2);
      a4 = div_const<15, 8>(yp2 - 4.0 * (yp1 + ym1) + 6.0 * y0 + ym2);
    }
    const double x1_ = 1.0 - 2.0 * cl;
    const double x2 = x1_ * x1_;
    const double x3 = x1_ * x2;
    const double ymin = dmin(y0, yp1);
    const double ymax = dmax(y0, yp1);
    double fmim = dmax(0.0, a0 * cl - a1 * (1.0 - x2) + a2 * (1.0 - x3) - a3 * (1.0 - x1_ * x3) + a4 * (1.0 - x2 * x3));
    fmim = dmin(fmim, y0 - ymin + fm_up);
    fmim = dmax(fmim, y0 - ymax + fm_up);
    fmim = dmax(0.0, fmim - (cl - clm) * y0);
    const double w = div_pos(y0, dmax(fmim + 1.e-15, y0));
    const double fm_dn = fmim * w;                                            // fm(i-1)
    st(i, y0 - fm_dn + fm_up);
    fm_up = fm_dn;
    yp2 = yp1; yp1 = y0; y0 = ym1; ym1 = ym2; ym2 = next;
  };
  int i = nf - 2;
  double n0 = yv(i - 3), n1 = yv(i - 4), n2 = yv(i - 5), n3 = yv(i - 6);
  while (i >= 1) {
    const double c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    n0 = yv(i - 7); n1 = yv(i - 8); n2 = yv(i - 9); n3 = yv(i - 10);          // the chunk after this one
    step(i, c0);
    if (i - 1 >= 1) step(i - 1, c1);
    if (i - 2 >= 1) step(i - 2, c2);
    if (i - 3 >= 1) step(i - 3, c3);
    i -= 4;
  }
  return y0 + fm_up;                                                          // y(1) = y(1) + fm(1)
}

// SUBROUTINE advsed0 (str.f90:5522-5579): upstream, levels 2 .. nf-1 updated in place from the bottom up;
// c(i) = -cneg(i).  psi(1) and psi(nf) are not changed by it.
template <class LD, class ST, class CNEG>
__device__ __forceinline__ void advsed0_pass(int nf, double x1, LD ld, ST st, CNEG cneg, double ytop)
{
  auto yu = [&](int j) -> double { return j <= nf - 2 ? ld(j) : (j == nf - 1 ? ytop : 0.0); };
  double ym = x1, yc = ld(1);
  const double c0 = -cneg(0);
  double fm_prev = -dmin(0.0, c0) * yc, fp_prev = dmax(0.0, c0) * ym;
  auto step = [&](int i, double yn) {
    const double ci = -cneg(i);
    const double fm_i = -dmin(0.0, ci) * yn, fp_i = dmax(0.0, ci) * yc;
    st(i, yc - fm_prev + fp_prev + fm_i - fp_i);
    fm_prev = fm_i; fp_prev = fp_i; ym = yc; yc = yn;
  };
  int i = 1;
  double n0 = yu(i + 1), n1 = yu(i + 2), n2 = yu(i + 3), n3 = yu(i + 4);
  while (i <= nf - 2) {
    const double c0_ = n0, c1 = n1, c2 = n2, c3 = n3;
    n0 = yu(i + 5); n1 = yu(i + 6); n2 = yu(i + 7); n3 = yu(i + 8);
    step(i, c0_);
    if (i + 1 <= nf - 2) step(i + 1, c1);
    if (i + 2 <= nf - 2) step(i + 2, c2);
    if (i + 3 <= nf - 2) step(i + 3, c3);
    i += 4;
  }
}

// ---- sedp -------------------------------------------------------------------------------------------------------
// Real spectra populate a few per cent of the 2-D grid, and only those classes settle (column sum > 1e-6,
// str.f90:2352).  Five launches per batch of columns:
//   sedp_scan_kernel   streams ff once (coalesced, HBM-bound): column sum of every class -> flag
//   sedp_list_kernel   per column: the settling classes, listed with the water index jt outermost, so that
//                      neighbours in the list have similar radii, i.e. the same advection scheme and similar numbers
//                      of sub-steps (the cost of a class grows with its terminal velocity)
//   sedp_units_kernel  work units = 64 consecutive list entries of one column; exclusive scan of the units per column
//   sedp_work_kernel   persistent blocks fetch units from a counter; one thread per settling class
//   sedp_diag_kernel   the diagnostics in the reference's class order
constexpr int SEDP_W = 64;       // classes per work unit = threads of the work kernel

__global__ void __launch_bounds__(256)
sedp_scan_kernel(mistra_sedp_args a, double *__restrict__ x0c, unsigned char *__restrict__ flag)
{
  __shared__ double s_detw[SED_MAXN];
  const int n = a.n, nf = a.nf, row = a.nka * a.nkt;
  for (int k = threadIdx.x; k < nf; k += 256) s_detw[k] = a.detw[k];
  __syncthreads();
  const size_t col = blockIdx.y;
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= row) return;
  const double *f = a.ff + col * (size_t)n * row + q;
  double xsum = 0.0;
#pragma unroll 11
  for (int k = 1; k < nf; ++k) xsum = xsum + f[(size_t)k * row] * s_detw[k];     // str.f90:2346-2349, in level order
  flag[col * row + q] = xsum > 1.e-6 ? 1 : 0;
  x0c[col * row + q] = 0.0;
}

__global__ void __launch_bounds__(256)
sedp_list_kernel(int nka, int nkt, const unsigned char *__restrict__ flag, unsigned short *__restrict__ list,
                 int *__restrict__ cnt, int *__restrict__ units)
{
  __shared__ int s_w[8], s_base;
  const int row = nka * nkt, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t col = blockIdx.x;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int p0 = 0; p0 < row; p0 += 256) {
    const int p = p0 + threadIdx.x;
    const int jt = p / nka, ia = p - jt * nka, q = ia * nkt + jt;
    const bool on = p < row && flag[col * row + q] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (on) list[col * row + off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)q;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_w[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { cnt[col] = s_base; units[col] = (s_base + SEDP_W - 1) / SEDP_W; }
}

__global__ void __launch_bounds__(1024)
sedp_units_kernel(int nc, const int *__restrict__ units, int *__restrict__ unit_off, int *__restrict__ counter)
{
  __shared__ int s_w[32], s_run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_run = 0; *counter = 0; }
  __syncthreads();
  for (int base = 0; base < nc; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nc ? units[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    int off = s_run;
    for (int w = 0; w < warp; ++w) off += s_w[w];
    if (i < nc) unit_off[i] = off + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_run = off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) unit_off[nc] = s_run;
}

__global__ void __launch_bounds__(SEDP_W, SEDP_MINB)
sedp_work_kernel(mistra_sedp_args a, int nc, const int *__restrict__ unit_off, const unsigned short *__restrict__ list,
                 const int *__restrict__ cnt, double *__restrict__ x0c, int *__restrict__ counter)
{
  extern __shared__ double sm[];
  __shared__ int s_u;
  const int n = a.n, nf = a.nf, row = a.nka * a.nkt;
  double *s_detw = sm, *s_deta = sm + nf, *s_t = sm + 2 * nf, *s_p = sm + 3 * nf, *s_rho = sm + 4 * nf,
         *s_eta = sm + 5 * nf;
  const int total = unit_off[nc];
  for (;;) {
    __syncthreads();                                                          // the unit before is done with the tables
    if (threadIdx.x == 0) s_u = atomicAdd(counter, 1);
    __syncthreads();
    const int u = s_u;
    if (u >= total) break;
    int lo = 0, hi = nc;                                                      // last column with unit_off[col] <= u
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (unit_off[mid] <= u) lo = mid; else hi = mid;
    }
    const size_t col = lo;
    const int i = (u - unit_off[lo]) * SEDP_W + threadIdx.x;
    for (int k = threadIdx.x; k < nf; k += SEDP_W) {
      const double tk = a.t[col * n + k], pk = a.p[col * n + k];
      s_detw[k] = a.detw[k]; s_deta[k] = a.deta[k]; s_t[k] = tk; s_p[k] = pk;
      s_rho[k] = rho_air(tk, pk); s_eta[k] = eta_air(tk);
    }
    __syncthreads();
    if (i >= cnt[col]) continue;
    const int q = list[col * row + i];
    double *f = a.ff + col * (size_t)n * row + q;
    const double rqq = a.rq[q], aq = rqq * 1.e-6;
    const double ww = -1. * vterm_lev(aq, s_t[nf - 1], s_p[nf - 1], s_rho[nf - 1], s_eta[nf - 1]);
    const double x3 = -s_deta[1], vdq = a.vd[col * row + q];
    double ytop = f[(size_t)(nf - 1) * row] * s_detw[nf - 1];                 // psi(nf)
    double dt0 = a.dt, x0 = 0.0, last = 0.0;
    bool first = true, finished = false;                                      // array holds ff (first) or psi; ff written
    for (int it = 0; dt0 > 0.1 && it < MISTRA_SED_MAXSUB; ++it) {             // str.f90:2355-2385
      const double dtmax = dmin(dt0, x3 / (ww));
      double c2 = dtmax / s_deta[1] * (-1. * vterm_lev(aq, s_t[1], s_p[1], s_rho[1], s_eta[1]));
      c2 = dmin(c2, dtmax / s_deta[1] * vdq * (-1.));
      dt0 = dt0 - dtmax;
      const bool fin = !(dt0 > 0.1);                                          // the last pass writes ff = psi / detw (2388-2391)
      auto cneg = [&](int lev) -> double {
        if (lev <= 1) return -c2;
        return -(dtmax / s_deta[lev] * (-1. * vterm_lev(aq, s_t[lev], s_p[lev], s_rho[lev], s_eta[lev])));
      };
      auto ld = [&](int i) -> double { const double v = f[(size_t)i * row]; return first ? v * s_detw[i] : v; };
      auto st = [&](int i, double v) {
        if (fin) { v = div_pos(v, s_detw[i]); if (i == nf - 2) last = v; }
        f[(size_t)i * row] = v;
      };
      const double x1 = ld(1);
      double p1 = x1;
      if (rqq < 1.0) advsed0_pass(nf, x1, ld, st, cneg, ytop); else p1 = advsed1_pass(nf, x1, ld, st, cneg, ytop);
      x0 = x0 + p1 - x1;
      first = false;
      finished = fin;
    }
    if (!finished)                                                            // no pass at all (dt <= 0.1) or the pass limit
      for (int k = 1; k < nf - 1; ++k) {
        double v = f[(size_t)k * row];
        if (first) v = v * s_detw[k];
        last = div_pos(v, s_detw[k]);
        f[(size_t)k * row] = last;
      }
    f[(size_t)(nf - 1) * row] = last;                                         // ff(jt,ia,nf) = ff(jt,ia,nf-1)
    x0c[col * row + q] = x0;
  }
}

// str.f90:2393-2409: x2 per class with the x0 the reference would hold at that point (an empty class keeps the x0 of
// the last settling class before it), then the running sums in the reference's class order - three dependent chains
// (ajs; trdep; ds1 | ds2), one thread each in three different warps.  One block per column; dynamic shared memory:
// 2 * row doubles.
__global__ void __launch_bounds__(256)
sedp_diag_kernel(mistra_sedp_args a, const double *__restrict__ x0c, const unsigned char *__restrict__ flag)
{
  extern __shared__ double s_x2[];
  const int row = a.nka * a.nkt, nkt = a.nkt;
  double *s_d = s_x2 + row;
  const size_t col = blockIdx.x;
  const double *x0 = x0c + col * row;
  const unsigned char *fl = flag + col * row;
  const double detw2 = a.detw[1], dt = a.dt;
  // every thread owns a contiguous stretch of classes; x0 entering the stretch = that of the last settling class before it
  const int per = (row + 255) / 256, q0 = threadIdx.x * per, q1 = min(row, q0 + per);
  double carry = 0.0;
  for (int q = min(q0, row) - 1; q >= 0; --q)
    if (fl[q]) { carry = x0[q]; break; }
  for (int q = q0; q < q1; ++q) {
    if (fl[q]) carry = x0[q];
    const double x2 = carry * a.e[q % nkt] * detw2;
    s_x2[q] = x2;
    s_d[q] = div_pos(x2, dt);
  }
  __syncthreads();
  double *dg = a.diag + col * 4;
  if (threadIdx.x == 0) {
    double ajs = 0.0;
#pragma unroll 4
    for (int q = 0; q < row; ++q) ajs = ajs + s_d[q];
    dg[0] = ajs;
  } else if (threadIdx.x == 32) {
    double trdep = dg[1];
#pragma unroll 4
    for (int q = 0; q < row; ++q) trdep = trdep + s_x2[q];
    dg[1] = trdep;
  } else if (threadIdx.x == 64) {
    double ds1 = dg[2], ds2 = dg[3];
    for (int ia = 0; ia < a.nka; ++ia) {
      const int kwi = a.kw[ia];
      const double *x = s_x2 + ia * nkt;
      for (int jt = 0; jt < nkt; ++jt)
        if (jt + 1 <= kwi) ds1 = ds1 + x[jt]; else ds2 = ds2 + x[jt];
    }
    dg[2] = ds1; dg[3] = ds2;
  }
}

// ---- sedl -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEDL_T, SEDL_MINB)
sedl_kernel(mistra_sedl_args a, double *__restrict__ s, int jx, int nchunk)
{
  extern __shared__ double sm[];
  const int n = a.n, nf = a.nf, nkc = a.nkc;
  double *s_detw = sm, *s_cc = sm + nf;
  const size_t col = blockIdx.y;
  const int kc = blockIdx.x / nchunk, l = (blockIdx.x % nchunk) * SEDL_T + threadIdx.x;
  for (int k = threadIdx.x; k < nf; k += SEDL_T) {                            // str.f90:2691-2701
    s_detw[k] = a.detw[k];
    if (k >= 1) {
      const double tk = a.t[col * n + k], pk = a.p[col * n + k];
      const double xxx = 0.01;
      const double x4 = dmax(xxx, 1.e6 * a.rc[(col * n + k) * nkc + kc]);
      double cc = (-1.0 * vterm_lev(x4 * 1.e-6, tk, pk, rho_air(tk, pk), eta_air(tk))) / a.deta[k];
      cc = dmin(cc, -1.0 * a.vt[(col * n + k) * nkc + kc] / a.deta[k]);
      if (k == 1) cc = dmin(cc, -1.0 / a.deta[1] * a.vdm[col * nkc + kc]);
      s_cc[k] = cc;
    }
  }
  __syncthreads();
  if (l >= jx) return;
  const size_t row = (size_t)nkc * jx;
  double *sl = s + col * (size_t)n * row + (size_t)kc * jx + l;
  double ytop = sl[(size_t)(nf - 1) * row] * s_detw[nf - 1];                  // psi(nf): advected, never written back
  double dt0 = a.dt, x0 = 0.0;
  const double xxxt = -.999 / s_cc[1];
  bool first = true, finished = false;
  for (int it = 0; dt0 > 0.1 && it < MISTRA_SED_MAXSUB; ++it) {               // 2708-2720
    const double dtmax = dmin(dt0, xxxt);
    dt0 = dt0 - dtmax;
    const bool fin = !(dt0 > 0.1);                                            // the last pass writes sl1 = psi / detw (2723-2725)
    auto ld = [&](int i) -> double { const double v = sl[(size_t)i * row]; return first ? v * s_detw[i] : v; };
    auto st = [&](int i, double v) { sl[(size_t)i * row] = fin ? div_pos(v, s_detw[i]) : v; };
    const double x1 = ld(1);
    const double p1 = advsed1_pass(nf, x1, ld, st, [&](int i) -> double { return -(s_cc[i < 1 ? 1 : i] * dtmax); }, ytop);
    x0 = x0 + p1 - x1;
    first = false;
    finished = fin;
  }
  if (!finished)
    for (int k = 1; k < nf - 1; ++k) {
      double v = sl[(size_t)k * row];
      if (first) v = v * s_detw[k];
      sl[(size_t)k * row] = div_pos(v, s_detw[k]);
    }
  sl[0] = sl[0] + x0 * a.deta[1];                                             // wet deposition, 2727
}

// ---- sedc -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sedc_kernel(mistra_sedc_args a, long long ncol)
{
  const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
  if (i >= ncol * a.j1) return;
  const long long col = i / a.j1;
  const int j = (int)(i % a.j1);
  const double Avogadro = 6.022140857e+23, x4 = 1.0;
  double *lev1 = a.s1 + col * (long long)a.n * a.j1 + j, *lev2 = lev1 + a.j1;
  const double w = a.vg[j], deta2 = a.deta[1];
  double s2 = *lev2;
  if (w >= 1.e-5) {                                                           // str.f90:2591-2594
    const double s12old = s2;
    s2 = s2 * exp(-a.dt / deta2 * w);
    *lev1 = *lev1 + (s12old - s2) * deta2;
  }
  *lev2 = s2 + a.es1[j] * x4 * a.dt * 1.e+4 / (a.detw[1] * Avogadro);          // 2596
}

// ---- self-test of div_const against the IEEE division -------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <int D, int K>
__device__ __forceinline__ int divc_bad(double x)
{
  const double a = div_const<D, K>(x), b = x / ((double)D * (double)(1 << K));
  return (__double_as_longlong(a) != __double_as_longlong(b) && !(a == 0.0 && b == 0.0) && !(a != a && b != b)) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
divc_selftest_kernel(long long n, unsigned long long seed, unsigned long long *__restrict__ bad)
{
  int mine = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const unsigned long long h = mix64(seed + (unsigned long long)i), g = mix64(h);
    // sign | exponent (every value incl. denormals, Inf, NaN in one sample of 16; else 2^-60 .. 2^60) | significand
    // (random; in one sample of 8 a multiple of the divisors with few trailing bits, the exactly representable quotients;
    // in one of 8 all ones or all zeros but the last bits)
    unsigned long long ex = (g & 15u) == 0 ? ((g >> 8) & 0x7ffull) : 963ull + ((g >> 8) % 121ull);
    unsigned long long man = h & 0xFFFFFFFFFFFFFull;
    const unsigned sel = (unsigned)(g >> 24) & 7u;
    if (sel == 0) man = ((h >> 20) * 15ull) & 0xFFFFFFFFFFFFFull;
    if (sel == 1) man = (g >> 32) & 1 ? 0xFFFFFFFFFFFFFull - (h & 7ull) : (h & 7ull);
    const double x = __longlong_as_double((long long)(((g >> 63) << 63) | (ex << 52) | man));
    mine += divc_bad<3, 3>(x) + divc_bad<3, 4>(x) + divc_bad<15, 7>(x) + divc_bad<3, 7>(x) + divc_bad<3, 8>(x) + divc_bad<15, 8>(x);
  }
  if (mine) atomicAdd(bad, (unsigned long long)mine);
}

std::recursive_mutex g_mu;
std::atomic<long long> g_launches{0};
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_stage[16], g_work[16];
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

int grow(Scratch &sc, size_t bytes)
{
  if (sc.bytes >= bytes) return 0;
  if (sc.p) { CKW(cudaDeviceSynchronize()); cudaFree(sc.p); sc.p = nullptr; sc.bytes = 0; }
  CKW(cudaMalloc(&sc.p, bytes));
  sc.bytes = bytes;
  return 0;
}

int current_device(int *dev)
{
  CKW(cudaGetDevice(dev));
  if (*dev < 0 || *dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  if (!g_attr[*dev]) {
    CKW(cudaFuncSetAttribute(sedp_work_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CKW(cudaFuncSetAttribute(sedl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CKW(cudaFuncSetAttribute(sedp_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    g_attr[*dev] = true;
  }
  return 0;
}

// host-buffer entries: stage `items` into one scratch block, run, copy the outputs back
struct Item { const void *h; size_t bytes; bool out; void **slot; };
template <class F>
int staged(std::vector<Item> &items, cudaStream_t st, F run)
{
  int dev = -1;
  if (int rc = current_device(&dev)) return rc;
  size_t total = 0;
  for (auto &it : items) total += (it.bytes + 255) & ~(size_t)255;
  if (int rc = grow(g_stage[dev], total)) return rc;
  char *p = g_stage[dev].p;
  for (auto &it : items) {
    if (!it.h) { *it.slot = nullptr; continue; }
    *it.slot = p;
    CKW(cudaMemcpyAsync(p, it.h, it.bytes, cudaMemcpyHostToDevice, st));
    p += (it.bytes + 255) & ~(size_t)255;
  }
  int rc = run();
  if (rc) { cudaStreamSynchronize(st); return rc; }
  for (auto &it : items)
    if (it.out && it.h) CKW(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKW(cudaStreamSynchronize(st));
  return 0;
}

bool bad_levels(int n, int nf) { return n < 6 || n > SED_MAXN || nf < 6 || nf > n; }

}  // namespace

extern "C" {

int mistra_sedp_device(int64_t ncol, const mistra_sedp_args *d_a, void *stream)
{
  if (ncol < 0 || !d_a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  const mistra_sedp_args &a = *d_a;
  if (bad_levels(a.n, a.nf) || a.nka < 1 || a.nkt < 1 || (int64_t)a.nka * a.nkt > 16000)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (6 <= nf <= n <= 512, 1 <= nka * nkt <= 16000)");
  if (!a.detw || !a.deta || !a.rq || !a.e || !a.kw || (ncol > 0 && (!a.t || !a.p || !a.vd || !a.ff || !a.diag)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int dev = -1;
  if (int rc = current_device(&dev)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t row = (size_t)a.nka * a.nkt;
  const size_t chunk = 32768;                                                 // columns per batch (gridDim.y)
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t ncm = std::min((size_t)ncol, chunk);
  const size_t o_flag = up(ncm * row * 8), o_list = o_flag + up(ncm * row), o_cnt = o_list + up(ncm * row * 2),
               o_units = o_cnt + up(ncm * 4), o_off = o_units + up(ncm * 4), o_ctr = o_off + up((ncm + 1) * 4);
  if (int rc = grow(g_work[dev], o_ctr + 256)) return rc;
  char *wb = g_work[dev].p;
  double *x0c = (double *)wb;
  unsigned char *flag = (unsigned char *)(wb + o_flag);
  unsigned short *list = (unsigned short *)(wb + o_list);
  int *cnt = (int *)(wb + o_cnt), *units = (int *)(wb + o_units), *unit_off = (int *)(wb + o_off), *counter = (int *)(wb + o_ctr);
  const size_t smem = (size_t)6 * a.nf * sizeof(double);
  static int nsm[16] = {};
  if (!nsm[dev]) CKW(cudaDeviceGetAttribute(&nsm[dev], cudaDevAttrMultiProcessorCount, dev));
  int per_sm = 1;
  CKW(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sedp_work_kernel, SEDP_W, smem));
  const unsigned workers = (unsigned)(nsm[dev] * std::max(per_sm, 1));
  for (size_t c0 = 0; c0 < (size_t)ncol; c0 += chunk) {
    const size_t nc = std::min(chunk, (size_t)ncol - c0);
    mistra_sedp_args b = a;
    b.t += c0 * a.n; b.p += c0 * a.n; b.vd += c0 * row; b.ff += c0 * a.n * row; b.diag += c0 * 4;
    sedp_scan_kernel<<<dim3((unsigned)((row + 255) / 256), (unsigned)nc), 256, 0, st>>>(b, x0c, flag);
    CKW(cudaGetLastError());
    sedp_list_kernel<<<(unsigned)nc, 256, 0, st>>>(a.nka, a.nkt, flag, list, cnt, units);
    CKW(cudaGetLastError());
    sedp_units_kernel<<<1, 1024, 0, st>>>((int)nc, units, unit_off, counter);
    CKW(cudaGetLastError());
    sedp_work_kernel<<<workers, SEDP_W, smem, st>>>(b, (int)nc, unit_off, list, cnt, x0c, counter);
    CKW(cudaGetLastError());
    sedp_diag_kernel<<<(unsigned)nc, 256, 2 * row * sizeof(double), st>>>(b, x0c, flag);
    CKW(cudaGetLastError());
    g_launches.fetch_add(5);
  }
  return 0;
}

int mistra_sedp(int64_t ncol, const mistra_sedp_args *a, void *stream)
{
  if (ncol < 0 || !a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  if (bad_levels(a->n, a->nf) || a->nka < 1 || a->nkt < 1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (6 <= nf <= n <= 512)");
  if (!a->detw || !a->deta || !a->rq || !a->e || !a->kw || (ncol > 0 && (!a->t || !a->p || !a->vd || !a->ff || !a->diag)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  const size_t nc = (size_t)ncol, n = a->n, row = (size_t)a->nka * a->nkt;
  mistra_sedp_args d = *a;
  std::vector<Item> items = {
      {a->detw, n * 8, false, (void **)&d.detw}, {a->deta, n * 8, false, (void **)&d.deta},
      {a->t, nc * n * 8, false, (void **)&d.t}, {a->p, nc * n * 8, false, (void **)&d.p},
      {a->rq, row * 8, false, (void **)&d.rq}, {a->e, (size_t)a->nkt * 8, false, (void **)&d.e},
      {a->kw, (size_t)a->nka * 4, false, (void **)&d.kw}, {a->vd, nc * row * 8, false, (void **)&d.vd},
      {a->ff, nc * n * row * 8, true, (void **)&d.ff}, {a->diag, nc * 4 * 8, true, (void **)&d.diag}};
  return staged(items, (cudaStream_t)stream, [&] { return mistra_sedp_device(ncol, &d, stream); });
}

int mistra_sedl_device(int64_t ncol, const mistra_sedl_args *d_a, void *stream)
{
  if (ncol < 0 || !d_a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  const mistra_sedl_args &a = *d_a;
  if (bad_levels(a.n, a.nf) || a.nkc < 1 || a.nkc_l < 0 || a.nkc_l > a.nkc || (a.sl1 && a.j2 < 1) || (a.sion1 && a.j6 < 1))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (6 <= nf <= n <= 512, 0 <= nkc_l <= nkc, j2, j6 >= 1)");
  if (!a.detw || !a.deta || (ncol > 0 && (!a.t || !a.p || !a.rc || !a.vt || !a.vdm)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0 || a.nkc_l == 0) return 0;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int dev = -1;
  if (int rc = current_device(&dev)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * a.nf * sizeof(double);
  const size_t chunk = 32768;
  for (size_t c0 = 0; c0 < (size_t)ncol; c0 += chunk) {
    const size_t nc = std::min(chunk, (size_t)ncol - c0);
    mistra_sedl_args b = a;
    b.t += c0 * a.n; b.p += c0 * a.n; b.rc += c0 * a.n * a.nkc; b.vt += c0 * a.n * a.nkc; b.vdm += c0 * a.nkc;
    for (int f = 0; f < 2; ++f) {
      double *s = f ? a.sion1 : a.sl1;
      const int jx = f ? a.j6 : a.j2;
      if (!s) continue;
      const int nchunk = (jx + SEDL_T - 1) / SEDL_T;
      sedl_kernel<<<dim3((unsigned)(nchunk * a.nkc_l), (unsigned)nc), SEDL_T, smem, st>>>(
          b, s + c0 * a.n * (size_t)a.nkc * jx, jx, nchunk);
      CKW(cudaGetLastError());
      g_launches.fetch_add(1);
    }
  }
  return 0;
}

int mistra_sedl(int64_t ncol, const mistra_sedl_args *a, void *stream)
{
  if (ncol < 0 || !a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  if (bad_levels(a->n, a->nf) || a->nkc < 1 || (a->sl1 && a->j2 < 1) || (a->sion1 && a->j6 < 1))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (6 <= nf <= n <= 512, j2, j6 >= 1)");
  if (!a->detw || !a->deta || (ncol > 0 && (!a->t || !a->p || !a->rc || !a->vt || !a->vdm)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  const size_t nc = (size_t)ncol, n = a->n, nkc = a->nkc;
  mistra_sedl_args d = *a;
  std::vector<Item> items = {
      {a->detw, n * 8, false, (void **)&d.detw}, {a->deta, n * 8, false, (void **)&d.deta},
      {a->t, nc * n * 8, false, (void **)&d.t}, {a->p, nc * n * 8, false, (void **)&d.p},
      {a->rc, nc * n * nkc * 8, false, (void **)&d.rc}, {a->vt, nc * n * nkc * 8, false, (void **)&d.vt},
      {a->vdm, nc * nkc * 8, false, (void **)&d.vdm},
      {a->sl1, a->sl1 ? nc * n * nkc * (size_t)a->j2 * 8 : 0, true, (void **)&d.sl1},
      {a->sion1, a->sion1 ? nc * n * nkc * (size_t)a->j6 * 8 : 0, true, (void **)&d.sion1}};
  return staged(items, (cudaStream_t)stream, [&] { return mistra_sedl_device(ncol, &d, stream); });
}

int mistra_sedc_device(int64_t ncol, const mistra_sedc_args *d_a, void *stream)
{
  if (ncol < 0 || !d_a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  const mistra_sedc_args &a = *d_a;
  if (a.n < 2 || a.j1 < 1) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (n >= 2, j1 >= 1)");
  if (!a.detw || !a.deta || !a.vg || !a.es1 || (ncol > 0 && !a.s1)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int dev = -1;
  if (int rc = current_device(&dev)) return rc;
  const long long tot = (long long)ncol * a.j1;
  sedc_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, (long long)ncol);
  CKW(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mistra_sedc(int64_t ncol, const mistra_sedc_args *a, void *stream)
{
  if (ncol < 0 || !a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  if (a->n < 2 || a->j1 < 1) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (n >= 2, j1 >= 1)");
  if (!a->detw || !a->deta || !a->vg || !a->es1 || (ncol > 0 && !a->s1)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (ncol == 0) return 0;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  const size_t nc = (size_t)ncol, n = a->n, j1 = a->j1;
  mistra_sedc_args d = *a;
  std::vector<Item> items = {
      {a->detw, n * 8, false, (void **)&d.detw}, {a->deta, n * 8, false, (void **)&d.deta},
      {a->vg, j1 * 8, false, (void **)&d.vg}, {a->es1, j1 * 8, false, (void **)&d.es1},
      {a->s1, nc * n * j1 * 8, true, (void **)&d.s1}};
  return staged(items, (cudaStream_t)stream, [&] { return mistra_sedc_device(ncol, &d, stream); });
}

int mistra_sed_divc_selftest(int64_t n, uint64_t seed, int64_t *mismatches)
{
  if (n < 0 || !mismatches) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int dev = -1;
  if (int rc = current_device(&dev)) return rc;
  if (int rc = grow(g_work[dev], 256)) return rc;
  unsigned long long *bad = (unsigned long long *)g_work[dev].p;
  CKW(cudaMemset(bad, 0, sizeof(*bad)));
  divc_selftest_kernel<<<148 * 8, 256>>>((long long)n, (unsigned long long)seed, bad);
  CKW(cudaGetLastError());
  unsigned long long h = 0;
  CKW(cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost));
  *mismatches = (int64_t)h;
  return 0;
}

int64_t mistra_sed_launch_count(void) { return g_launches.load(); }

}  // extern "C"
