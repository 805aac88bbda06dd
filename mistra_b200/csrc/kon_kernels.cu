// Condensation / evaporation on the 2-D particle grid (include/mistra_kon.h): CUDA kernel +
// C-ABI entries.  Role in the reference: SUBROUTINE subkon (/root/reference/src/str.f90:
// 4987-5204) with SUBROUTINE advec (5321-5516) and diff_wat_vap / therm_conduct_air / xl21 /
// p21 (5210-5315, 7640-7693), for every humid layer at once.
//
// Mapping: one CTA per layer.  Five nka x nkt tiles live in shared memory for the whole
// layer (196 kB for 70 x 70, one CTA per SM): the growth-rate coefficients cd, cr, sr (set
// up once by all threads, 18-band radiative sum from the L2-resident qabs table), the
// spectrum before the step falt (cp.async from HBM) and the advected spectrum ffk (written
// back once).  Per secant iteration the advection of advec is split in two: (A) one thread per
// grid point (ia, i) walks the Courant path of that bin and evaluates Bott's polynomial flux
// (the expensive, independent part: divisions, u(k) from cd, cr, sr on the fly) and parks
// the target bin and the flux x1 in shared memory; (B) one thread per target bin (ia, k) adds
// the contributions that land on y(k) in increasing i - the order of the reference's loop - so
// every y(k) sees the same additions in the same order and the result is the sequential one
// to the last bit (the scan is bounded by the row's largest bin displacement).  Rows go through A/B in chunks sized by the shared memory left.  The
// liquid-water change is summed per row, then over rows by one thread, which also runs the
// secant update.  Compiled without FMA contraction (build.py) so that the arithmetic is the
// reference's up to the CUDA exp/pow (<= 2 ulp) and the per-row summation of dwsum.
#include "../../include/mistra_kon.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int MB = MISTRA_MB;
constexpr int KON_THREADS = 512;
constexpr int KON_MAX_NKA = 128;
constexpr int MAXK = 128;

struct KonGridDev {
  int nka, nkt, ka;
  double a0m, dlne;
  const double *en, *rn, *b0m, *ew, *e, *dew, *rw, *qabs, *rq;
  const int *kw;
};

// Per-layer arrays.  Two views of the same kernel:
//  * subkon view (full == 0): the arguments of "call subkon" (str.f90:4705);
//  * kon view (full == 1): the COMMON arrays the layer loop of SUBROUTINE kon reads and
//    writes (str.f90:4615-4772), incl. the dry branch (equil) and, with chem, the bin sums.
struct KonArgs {
  int full, chem;
  double *ff;                                   // [ncell][nka][nkt] in/out
  const double *totr;                           // [ncell][18]
  const double *pp;                             // [ncell]
  const int *kr;                                // [ncell] nar(k)
  int *status;                                  // [ncell] or null
  // subkon view
  const double *dfdt, *feualt, *tn, *xm1n;      // in
  double *to, *xm1o;                            // in/out
  // kon view
  double *t, *talt, *xm1, *xm1a, *feu, *dfddt, *xm2, *dtcon;
  double *vol1_a, *vol1_d, *part_o_a, *part_o_d, *part_n_a, *part_n_d;   // [ncell][nka] or null
  double *vol2, *pntot;                                                  // [ncell][4]   or null
};

// FUNCTION rgl (str.f90:2164-2251): equilibrium radius of a solution droplet at relative
// humidity feu < 1 by Newton iteration on the Koehler equation.
__device__ double rgl(double r_dry, double a, double b, double feu)
{
  if (feu >= 1.0) return r_dry;
  const double zlogf = log(feu);
  const double alpha = a / r_dry;
  double xalt = exp(feu), xneu = xalt;
  for (int ij = 1; ij <= 100; ++ij) {
    const double x3 = xalt * xalt * xalt;
    const double falt = (x3 - 1.0) * (xalt * zlogf - alpha) + b * xalt;
    const double fstralt = (4.0 * x3 - 1.0) * zlogf - 3.0 * (xalt * xalt) * alpha + b;
    xneu = xalt - falt / fstralt;
    if (fabs(xneu - xalt) < 1.e-7 * xalt) break;
    xalt = xneu;
  }
  return r_dry * xneu;
}

// constants.f90:48-83
__device__ constexpr double kGasConst = 8.3144743, kMair = 28.96546e-3, kMwat = 18.01528e-3;
__device__ constexpr double kCp = 1005.0, kRhow = 1000.0, kPi = 3.1415926535897932;

__device__ __forceinline__ double xl21(double t) { return 3138708. + (-2339.4) * t; }                       // str.f90:7661
__device__ __forceinline__ double p21(double t) { return 610.7 * exp(17.15 * (t - 273.15) / (t - 38.33)); } // str.f90:7691
__device__ __forceinline__ double therm_conduct_air(double t) { return 4.39e-3 + 7.1e-5 * t; }              // str.f90:5313
__device__ __forceinline__ double diff_wat_vap(double t, double p)                                          // str.f90:5255-5257
{
  const double cst2 = 0.211e-4 * 101325. / 53286.64010226011;  // cst*P0/T0**exponent, a compile-time constant there
  return cst2 * pow(t, 1.94) / p;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

struct RowCoef {
  const double *cd, *cr, *sr;  // this row's coefficients in shared memory
  double fquer, dlne;
  int nkt;
  // c(jt) of str.f90:5167 and the upstream velocities u(jt) of str.f90:5170-5175, 1-based
  __device__ __forceinline__ double c(int jt) const
  {
    return (cd[jt - 1] * (fquer - sr[jt - 1]) - cr[jt - 1]) / dlne;
  }
  __device__ __forceinline__ double u(int jt) const
  {
    if (jt == 1) return fmax(0.0, c(1));
    if (jt == nkt) return fmin(0.0, c(nkt - 1));
    const double a = c(jt), b = c(jt - 1);
    return 0.5 * (a + fabs(a) + b - fabs(b));
  }
};

// SUBROUTINE advec (str.f90:5321-5516), part A for one source bin i (1-based) of a row:
// z = spectrum before (read only).  Returns the kind of contribution in the high byte and
// the target bin (1-based) in the low byte; *x1 = flux into k_low+1 for ADV_SPLIT.
enum { ADV_SKIP = 0, ADV_ADD = 1, ADV_SPLIT = 2, ADV_ERR = 3 };
__device__ __forceinline__ unsigned advec_flux(int nkt, double dt, const RowCoef &rc, const double *__restrict__ z,
                                               int i, double *x1_out)
{
  const double ymin = 1.e-32;
  const double zi = z[i - 1];
  if (zi < ymin) return ADV_SKIP << 8;
  int k2 = 0, k1, k = i;
  double dt0, dt1 = dt, x0;
  double uk = rc.u(k);
  if (fabs(uk) > 0.0) dt0 = fmin(1.0 / fabs(uk), dt1);
  else return (ADV_ADD << 8) | (unsigned)k;
  x0 = (double)k + uk * dt0;
  dt1 = dt1 - dt0;
  k1 = k;
  while (dt1 > 1.e-7) {
    if (uk < 0.0) k = k - 1; else k = k + 1;
    if (k == k2) return (ADV_ADD << 8) | (unsigned)k;
    k2 = k1;
    k1 = k;
    if (k < 1 || k > nkt) return ADV_ERR << 8;
    uk = rc.u(k);
    if (fabs(uk) > 0.0) dt0 = fmin(1.0 / fabs(uk), dt1);
    else return (ADV_ADD << 8) | (unsigned)k;
    x0 = (double)k + uk * dt0;
    dt1 = dt1 - dt0;
  }
  const int k_low = (int)floor(x0), k_high = k_low + 1;
  const double c0 = x0 - (double)k_low;
  if (k_low < 1 || (k_high > nkt && c0 > 0.0)) return ADV_ERR << 8;
  if (!(c0 > 0.0)) return (ADV_ADD << 8) | (unsigned)k_low;
  double x1;
  if (i == 1 || i == nkt) {
    x1 = c0 * zi;
  } else if (i == 2 || i == nkt - 1) {
    const double al = 1.0 - 2.0 * c0, al2 = al * al;
    const double a0 = (26.0 * zi - z[i] - z[i - 2]) / 24.0;
    const double a1 = (z[i] - z[i - 2]) / 16.0;
    const double a2 = (z[i] + z[i - 2] - 2.0 * zi) / 48.0;
    x1 = fmin(zi, a0 * c0 + a1 * (1.0 - al2) + a2 * (1.0 - al2 * al));
  } else {
    const double al = 1.0 - 2.0 * c0, al2 = al * al, al3 = al2 * al;
    const double zp2 = z[i + 1], zp1 = z[i], zm1 = z[i - 2], zm2 = z[i - 3];
    const double a0 = (9.0 * (zp2 + zm2) - 116.0 * (zp1 + zm1) + 2134.0 * zi) / 1920.0;
    const double a1 = (-5.0 * (zp2 - zm2) + 34.0 * (zp1 - zm1)) / 384.0;
    const double a2 = (-zp2 + 12.0 * (zp1 + zm1) - 22.0 * zi - zm2) / 384.0;
    const double a3 = (zp2 - 2.0 * (zp1 - zm1) - zm2) / 768.0;
    const double a4 = (zp2 - 4.0 * (zp1 + zm1) + 6.0 * zi + zm2) / 3840.0;
    x1 = fmin(zi, a0 * c0 + a1 * (1.0 - al2) + a2 * (1.0 - al3) + a3 * (1.0 - al2 * al2)
                      + a4 * (1.0 - al2 * al3));
  }
  *x1_out = fmax(0.0, x1);
  return (ADV_SPLIT << 8) | (unsigned)k_low;
}

// Part B for one target bin k (1-based) of a row: the additions advec's loop over i makes to
// y(k), in its order (increasing i).  Source bin i contributes to y(k_low) and, when split, to
// y(k_low+1); `w` bounds |k_low - i| over the row, so only i in [k-1-w, k+w] can reach k;
// `iend` = first source bin at which the reference aborts (nkt+1 if none): the sequential loop
// never gets to the bins from there on.  y starts from 0 as in advec.
__device__ __forceinline__ double advec_gather(int nkt, int k, int w, int iend, const double *__restrict__ z,
                                               const double *__restrict__ x1, const unsigned short *__restrict__ code)
{
  double y = 0.0;
  const int lim = min(nkt, iend - 1);
  if (w <= 1) {
    // usual case (Courant numbers below one bin): only i = k-2 .. k+1 can reach k; the four
    // candidates are fetched at once, the additions keep their order
    unsigned cd[4];
    double zz[4], ff[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int i = k - 2 + d;
      const bool ok = (i >= 1) && (i <= lim);
      cd[d] = ok ? (unsigned)code[ok ? i - 1 : 0] : (unsigned)(ADV_SKIP << 8);
      zz[d] = z[ok ? i - 1 : 0];
      ff[d] = x1[ok ? i - 1 : 0];
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const unsigned kind = cd[d] >> 8;
      const int kl = (int)(cd[d] & 255u);
      if (kind == ADV_ADD) {
        if (kl == k) y = y + zz[d];
      } else if (kind == ADV_SPLIT) {
        if (kl == k) y = y + zz[d] - ff[d];
        else if (kl == k - 1) y = y + ff[d];
      }
    }
    return y;
  }
  const int lo = max(1, k - 1 - w), hi = min(k + w, lim);
  for (int i = lo; i <= hi; ++i) {
    const unsigned cd = code[i - 1], kind = cd >> 8;
    const int kl = (int)(cd & 255u);
    if (kind == ADV_ADD) {
      if (kl == k) y = y + z[i - 1];
    } else if (kind == ADV_SPLIT) {
      if (kl == k) y = y + z[i - 1] - x1[i - 1];
      else if (kl == k - 1) y = y + x1[i - 1];
    }
  }
  return y;
}

#ifdef KON_PROF
__device__ unsigned long long g_kon_prof[8];
#define KP_START() long long kp_prev = clock64()
#define KP(i) { const long long kp_now = clock64(); if (threadIdx.x == 0) atomicAdd(&g_kon_prof[i], (unsigned long long)(kp_now - kp_prev)); kp_prev = kp_now; }
#else
#define KP_START()
#define KP(i)
#endif

__global__ void __launch_bounds__(KON_THREADS)
kon_subkon_kernel(KonGridDev g, long long ncell, double dt, KonArgs A, int rch)
{
  double *__restrict__ ffk_all = A.ff;
  const double *__restrict__ totr_all = A.totr;
  int *__restrict__ status = A.status;
  extern __shared__ __align__(16) double sm[];
  const int nka = g.nka, nkt = g.nkt, ntile = nka * nkt;
  double *s_cd = sm, *s_cr = s_cd + ntile, *s_sr = s_cr + ntile, *s_falt = s_sr + ntile, *s_ffk = s_falt + ntile;
  double *s_dw = s_ffk + ntile;      // [nka] liquid-water change per dry class
  double *s_e = s_dw + nka;          // [nkt]
  double *s_totr = s_e + nkt;        // [MB]
  double *s_it = s_totr + MB;        // [2] fquer, spare
  int *s_flag = (int *)(s_it + 2);   // [0] stop code, [1] advec error, [2..2+nka) kr0 switch per class
  double *s_x1 = (double *)(s_flag + ((nka + 5) & ~1));          // [rch][nkt] flux of part A
  unsigned short *s_code = (unsigned short *)(s_x1 + rch * nkt);   // [rch][nkt] kind << 8 | target bin
  int *s_w = (int *)(s_code + ((rch * nkt + 1) & ~1));             // [rch] largest |k_low - i| of the row
  int *s_iend = s_w + rch;                                         // [rch] first aborting source bin
  const double r0 = kGasConst / kMair, r1 = kGasConst / kMwat;

  for (int i = threadIdx.x; i < nkt; i += blockDim.x) s_e[i] = g.e[i];
  // the reference lowers kr0 from 3 to 2 at the first class with rn < 0.5 and never
  // restores it (str.f90:5141): a prefix flag over ia
  if (threadIdx.x == 0) {
    int seen = 0;
    for (int ia = 0; ia < nka; ++ia) { seen |= (g.rn[ia] < 0.5); s_flag[2 + ia] = seen; }
  }

  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    __syncthreads();
    KP_START();
    double *ffk = ffk_all + (size_t)c * ntile;
    if ((ntile & 1) == 0)
      for (int i = threadIdx.x; i < (ntile >> 1); i += blockDim.x) cp_async16(s_falt + 2 * i, ffk + 2 * i);
    else
      for (int i = threadIdx.x; i < ntile; i += blockDim.x) cp_async8(s_falt + i, ffk + i);
    if (threadIdx.x < MB) s_totr[threadIdx.x] = totr_all[c * MB + threadIdx.x];
    // ---- layer scalars (str.f90:5104-5127), by every thread ----
    // subkon's arguments; in the kon view they come from the COMMON arrays (str.f90:4676-4703)
    const double to0 = A.full ? A.talt[c] : A.to[c], xm1o0 = A.full ? A.xm1a[c] : A.xm1o[c];
    const double tn = A.full ? A.t[c] : A.tn[c], xm1n = A.full ? A.xm1[c] : A.xm1n[c], pp = A.pp[c];
    const double feualt = A.full ? A.feu[c] : A.feualt[c];
    const double dfdt_c = A.full ? A.dfddt[c] : A.dfdt[c];
    const bool dry = A.full && feualt < 0.7;          // str.f90:4663: Koehler equilibrium instead
    if (A.full) {
      // the tile is needed right away (bin sums, dry branch)
      cp_async_wait_all();
      __syncthreads();
      if (A.chem) {
        // particle number and volume per dry class before the step (str.f90:4625-4659)
        if (threadIdx.x < nka) {
          const int ia = threadIdx.x, kwa = g.kw[ia];
          const double z4pi3 = 4.0 * kPi / 3.0;
          double va = 0.0, pa = 0.0, vd = 0.0, pd = 0.0;
          for (int jt = 0; jt < nkt; ++jt) {
            const double f = s_falt[ia * nkt + jt], r = g.rq[ia * nkt + jt];
            const double v = f * z4pi3 * (r * r * r);
            if (jt < kwa) { va = va + v; pa = pa + f; } else { vd = vd + v; pd = pd + f; }
          }
          A.vol1_a[c * nka + ia] = va; A.part_o_a[c * nka + ia] = pa;
          A.vol1_d[c * nka + ia] = vd; A.part_o_d[c * nka + ia] = pd;
          s_cd[ia] = va; s_cd[nka + ia] = vd;        // scratch: the coefficient tile is not set up yet
        }
        __syncthreads();
        if (threadIdx.x < 4) {                       // vol2(1..4): sums over the classes of the bin
          const int kc = threadIdx.x, lo = (kc & 1) ? g.ka : 0, hi = (kc & 1) ? nka : g.ka;
          double sacc = 0.0;
          for (int ia = lo; ia < hi; ++ia) sacc = sacc + s_cd[(kc >= 2 ? nka : 0) + ia];
          A.vol2[c * 4 + kc] = sacc;
        }
        __syncthreads();
      }
    }
    if (dry) {
      // ---- SUBROUTINE equil, case 1 (str.f90:4801-4981) on this layer ----
      const double feun = xm1n * pp / ((0.62198 + 0.37802 * xm1n) * p21(tn));   // str.f90:4664
      const double a0e = g.a0m / tn;
      if (threadIdx.x < nka) {
        const int ia = threadIdx.x;
        double tot = 0.0;
        for (int jt = 0; jt < nkt; ++jt) { tot = tot + s_falt[ia * nkt + jt]; s_ffk[ia * nkt + jt] = 0.0; }
        const double rn = g.rn[ia];
        const double rg = rgl(rn, a0e, g.b0m[ia] * 2.0, feun);       // zrho_frac = rho3/rhow = 2
        const double eg = (4.e-09 * kPi / 3.0) * (rg * rg * rg - rn * rn * rn);
        int jt = 1;
        while (jt < nkt && eg > g.ew[jt - 1]) jt = jt + 1;          // (clamped at the last bin)
        s_ffk[ia * nkt + jt - 1] = tot;
        s_dw[ia] = tot * s_e[jt - 1];                               // this class' share of xm2
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double x2 = 0.0;
        for (int ia = 0; ia < nka; ++ia) x2 = x2 + s_dw[ia];
        A.feu[c] = feun;
        A.xm2[c] = x2;
        A.dtcon[c] = 0.0;
        if (status) status[c] = 0;
      }
    } else {
    const double zxl21 = xl21(to0);
    const double xldcp = zxl21 / kCp;
    const double xka = therm_conduct_air(to0);
    const double xdv = diff_wat_vap(to0, pp);
    const double xl = 24.483 * to0 / pp;
    const double deltav = 1.3 * xl, deltat = 2.7 * xl;
    const double rho = pp / (r0 * to0 * (1.0 + 0.61 * xm1o0));
    const double rho21 = p21(to0) / (r1 * to0);
    const double rho21s = (zxl21 / (r1 * to0) - 1.0) * rho21 / to0;
    const double a0 = g.a0m / to0;
    const double xdv0 = xdv * sqrt(2.0 * kPi / (r1 * to0)) / 3.6e-08;
    const double xka0 = xka * sqrt(2.0 * kPi / (r0 * to0)) / (7.e-07 * rho * kCp);
    const int kr = A.kr[c];
    __syncthreads();  // s_totr visible
    const int ib0 = (s_totr[0] < 1.0) ? 7 : 1;
    KP(0);
    // ---- growth-rate coefficients over the grid (str.f90:5128-5149) ----
    for (int q = threadIdx.x; q < ntile; q += blockDim.x) {
      const int ia = q / nkt + 1, jt = q - (ia - 1) * nkt + 1;
      const int jtp = jt + 1 < nkt ? jt + 1 : nkt;
      const double de0 = g.dew[jt - 1], dep = g.dew[jtp - 1], de0p = de0 + dep;
      const double rk = g.rw[q];
      const double srq = fmax(0.1, exp(a0 / rk - g.b0m[ia - 1] * g.en[ia - 1] / g.ew[jt - 1]));
      const double xdvs = xdv / (rk / (rk + deltav) + xdv0 / rk);
      const double xkas = xka / (rk / (rk + deltat) + xka0 / rk);
      const double x1 = kRhow * (zxl21 + xkas / (xdvs * rho21s * srq));
      const int kr0 = (kr == 3 && s_flag[2 + ia - 1]) ? 2 : kr;
      // device table: [kr][ia][ib][jt] of qabs(ib,jt)*de0 + qabs(ib,jtp)*dep (built at upload)
      const double *qa = g.qabs + ((size_t)(kr0 - 1) * nka + (ia - 1)) * nkt * MB;
      double qv[MB];
#pragma unroll
      for (int ib = 1; ib <= MB; ++ib) qv[ib - 1] = qa[(ib - 1) * nkt + jt - 1];
      double rad = 0.0;
#pragma unroll
      for (int ib = 1; ib <= MB; ++ib)
        if (ib >= ib0) rad = rad + s_totr[ib - 1] * qv[ib - 1] / de0p;
      s_sr[q] = srq;
      s_cd[q] = 3.e12 * rho21 * xkas / (x1 * rk * rk * rho21s * srq);
      s_cr[q] = rad * 7.5e5 / (rk * x1) - kRhow * 4190. * (tn - to0) / (dt * x1);
    }
    // ---- secant iteration on the mean saturation ratio (str.f90:5151-5201) ----
    double feuneu = feualt + dfdt_c * dt;
    if (feualt < 0.95) feuneu = xm1n * pp / (p21(tn) * (.62198 + .37802 * xm1n));
    double fquer = 0.5 * (feuneu + feualt);
    double res = 0.0, fqa = 0.0, to = to0, xm1o = xm1o0;
    const double aa0 = 1.0 / dt;
    int st = -1;
    if (threadIdx.x == 0) { s_flag[0] = 0; s_flag[1] = 0; }
    cp_async_wait_all();
    __syncthreads();
    KP(1);
    for (int itk = 1; itk <= 10; ++itk) {
      for (int r0 = 0; r0 < nka; r0 += rch) {
        const int nr = min(rch, nka - r0);
        if (threadIdx.x < nr) { s_w[threadIdx.x] = 0; s_iend[threadIdx.x] = nkt + 1; }
        __syncthreads();
        // part A: one thread per (dry class, source bin)
        for (int q = threadIdx.x; q < nr * nkt; q += blockDim.x) {
          const int rr = q / nkt, i = q - rr * nkt + 1, r = (r0 + rr) * nkt;
          RowCoef rc{s_cd + r, s_cr + r, s_sr + r, fquer, g.dlne, nkt};
          double x1 = 0.0;
          const unsigned cd = advec_flux(nkt, dt, rc, s_falt + r, i, &x1);
          s_code[q] = (unsigned short)cd;
          s_x1[q] = x1;
          const unsigned kind = cd >> 8;
          if (kind == ADV_ERR) { atomicMin(&s_iend[rr], i); s_flag[1] = 1; }
          else if (kind != ADV_SKIP) {
            const int d = abs((int)(cd & 255u) - i);
            if (d > s_w[rr]) atomicMax(&s_w[rr], d);
          }
        }
        __syncthreads();
        KP(2);
        // part B: one thread per (dry class, target bin), ordered additions
        for (int q = threadIdx.x; q < nr * nkt; q += blockDim.x) {
          const int rr = q / nkt, k = q - rr * nkt + 1, r = (r0 + rr) * nkt;
          s_ffk[r + k - 1] = advec_gather(nkt, k, s_w[rr], s_iend[rr], s_falt + r, s_x1 + rr * nkt, s_code + rr * nkt);
        }
        if (r0 + rch < nka) __syncthreads();   // the flux buffer, s_w and s_iend are reused by the next chunk
        KP(3);
      }
      __syncthreads();
      // each class' water change (per row, in bin order; the differences and products do not
      // depend on the running sum, so they are formed four bins ahead of it)
      if (threadIdx.x < nka) {
        const int r = threadIdx.x * nkt;
        double dw = 0.0;
        int jt = 0;
        for (; jt + 4 <= nkt; jt += 4) {
          const double p0 = (s_ffk[r + jt] - s_falt[r + jt]) * s_e[jt];
          const double p1 = (s_ffk[r + jt + 1] - s_falt[r + jt + 1]) * s_e[jt + 1];
          const double p2 = (s_ffk[r + jt + 2] - s_falt[r + jt + 2]) * s_e[jt + 2];
          const double p3 = (s_ffk[r + jt + 3] - s_falt[r + jt + 3]) * s_e[jt + 3];
          dw = dw + p0; dw = dw + p1; dw = dw + p2; dw = dw + p3;
        }
        for (; jt < nkt; ++jt) dw = dw + (s_ffk[r + jt] - s_falt[r + jt]) * s_e[jt];
        s_dw[threadIdx.x] = dw;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int stop = 0;
        if (s_flag[1]) { st = -2; stop = 1; }
        else {
          double dwsum = 0.0;
          for (int ia = 0; ia < nka; ++ia) dwsum = dwsum + s_dw[ia];
          const double dmsum = dwsum / rho;
          const double dtsum = xldcp * dmsum;
          xm1o = xm1n - dmsum;
          to = tn + dtsum;
          const double p1 = xm1o * pp / (0.62198 + 0.37802 * xm1o);
          feuneu = p1 / p21(to);
          const double resold = res;
          res = feuneu + feualt - 2.0 * fquer;
          if (fabs(res) < 1.e-6) { st = itk; stop = 1; }
          else {
            const double dres = res - resold;
            double aa = aa0;
            if (itk > 1 && fabs(dres) > 1.e-8) aa = (fqa - fquer) / dres;
            fqa = fquer;
            fquer = fquer + aa * res;
          }
        }
        s_it[0] = fquer;
        s_flag[0] = stop;
      }
      __syncthreads();
      KP(4);
      fquer = s_it[0];
      if (s_flag[0]) break;
    }
    // ---- write back of the humid branch (str.f90:4708-4721) ----
    if (threadIdx.x < nka && A.full) {               // xm2: liquid water, per class then over classes
      double x2 = 0.0;
      for (int jt = 0; jt < nkt; ++jt) x2 = x2 + s_ffk[threadIdx.x * nkt + jt] * s_e[jt];
      s_dw[threadIdx.x] = x2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (A.full) {
        A.t[c] = to; A.talt[c] = to; A.xm1[c] = xm1o; A.xm1a[c] = xm1o;
        const double fn = xm1o * pp / ((0.62198 + 0.37802 * xm1o) * p21(to));
        A.feu[c] = fn;
        A.dfddt[c] = (fn - feualt) / dt;
        double x2 = 0.0;
        for (int ia = 0; ia < nka; ++ia) x2 = x2 + s_dw[ia];
        A.xm2[c] = x2;
        A.dtcon[c] = (to - tn) / dt;
      } else {
        A.to[c] = to;
        A.xm1o[c] = xm1o;
      }
      if (status) status[c] = st;
    }
    }  // humid branch
    __syncthreads();
    if (A.full && A.chem) {
      // particle number per class and bin after the step (str.f90:4724-4770)
      if (threadIdx.x < nka) {
        const int ia = threadIdx.x, kwa = g.kw[ia];
        double pa = 0.0, pd = 0.0;
        for (int jt = 0; jt < nkt; ++jt) {
          const double f = s_ffk[ia * nkt + jt];
          if (jt < kwa) pa = pa + f; else pd = pd + f;
        }
        A.part_n_a[c * nka + ia] = pa; A.part_n_d[c * nka + ia] = pd;
        s_cd[ia] = pa; s_cd[nka + ia] = pd;
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        const int kc = threadIdx.x, lo = (kc & 1) ? g.ka : 0, hi = (kc & 1) ? nka : g.ka;
        double sacc = 0.0;
        for (int ia = lo; ia < hi; ++ia) sacc = sacc + s_cd[(kc >= 2 ? nka : 0) + ia];
        A.pntot[c * 4 + kc] = sacc;
      }
    }
    // the advected (or equilibrated) spectrum goes back once
    if ((ntile & 1) == 0) {
      const double2 *s2 = reinterpret_cast<const double2 *>(s_ffk);
      double2 *d2 = reinterpret_cast<double2 *>(ffk);
      for (int i = threadIdx.x; i < (ntile >> 1); i += blockDim.x) d2[i] = s2[i];
    } else {
      for (int i = threadIdx.x; i < ntile; i += blockDim.x) ffk[i] = s_ffk[i];
    }
    KP(5);
  }
}

// ---- host side ---------------------------------------------------------------------------
// recursive: the host-buffer wrappers hold it across their scratch handling and the launch (launch() takes it too)
std::recursive_mutex g_mu;
std::atomic<long long> g_launches{0};

struct GridCache {
  bool valid = false, attr_set = false;
  int nka = 0, nkt = 0, num_sm = 0;
  std::vector<double> host;   // concatenated copy for change detection
  double *d = nullptr;        // device copy, same concatenation
  size_t cap = 0;
  std::vector<int> kw;
  int *d_kw = nullptr;
};
GridCache g_cache[16];
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_scratch[16];

#define CKK(call)                                                                       \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return mistra_internal_fail(e_ == cudaErrorMemoryAllocation ? MISTRA_KPP_ENOMEM   \
                                  : (e_ == cudaErrorNoDevice ? MISTRA_KPP_ENODEVICE     \
                                                             : MISTRA_KPP_ECUDA),       \
                                  std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

int check_grid(const mistra_kon_grid *g)
{
  if (!g || !g->en || !g->rn || !g->b0m || !g->ew || !g->e || !g->dew || !g->rw || !g->qabs)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null grid");
  if (g->nka < 1 || g->nka > KON_MAX_NKA || g->nkt < 5 || g->nkt > MAXK)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "grid size out of range (nka <= 128, 5 <= nkt <= 128)");
  if (!(g->dlne > 0.0)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "dlne <= 0");
  return 0;
}

size_t smem_bytes(const mistra_kon_grid *g)
{
  const size_t ntile = (size_t)g->nka * g->nkt;
  return sizeof(double) * (5 * ntile + g->nka + g->nkt + MB + 2) + sizeof(int) * ((g->nka + 5) & ~1);
}

// rows per part-A/B chunk: what the flux buffer (8 + 2 bytes per grid point) can hold
int rows_per_chunk(const mistra_kon_grid *g)
{
  const size_t base = smem_bytes(g), cap = 227 * 1024;
  if (base >= cap) return 0;
  size_t r = (cap - base - 32) / ((size_t)g->nkt * 10 + 8);
  return (int)(r > (size_t)g->nka ? (size_t)g->nka : r);
}

int grid_to_device(const mistra_kon_grid *g, cudaStream_t st, KonGridDev *out, GridCache **cache, bool full)
{
  if (full) {
    if (!g->kw || !g->rq) return mistra_internal_fail(MISTRA_KPP_EINVAL, "mistra_kon_layers needs grid kw and rq");
    if (g->ka < 0 || g->ka > g->nka) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad ka");
    for (int i = 0; i < g->nka; ++i)
      if (g->kw[i] < 0 || g->kw[i] > g->nkt) return mistra_internal_fail(MISTRA_KPP_EINVAL, "kw out of range");
  }
  int dev = -1;
  CKK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  GridCache &gc = g_cache[dev];
  const size_t nka = g->nka, nkt = g->nkt;
  const size_t n_all = 3 * nka + 3 * nkt + 2 * nka * nkt + (size_t)MISTRA_JPTAERRAD * nka * nkt * MB;
  std::vector<double> h;
  h.reserve(n_all);
  auto put = [&](const double *p, size_t n) { h.insert(h.end(), p, p + n); };
  put(g->en, nka); put(g->rn, nka); put(g->b0m, nka); put(g->ew, nkt); put(g->e, nkt); put(g->dew, nkt);
  put(g->rw, nka * nkt); put(g->qabs, (size_t)MISTRA_JPTAERRAD * nka * nkt * MB);
  if (g->rq) put(g->rq, nka * nkt); else h.insert(h.end(), nka * nkt, 0.0);
  std::vector<int> kwv(nka, 0);
  if (g->kw) kwv.assign(g->kw, g->kw + nka);
  const bool same = gc.valid && gc.nka == g->nka && gc.nkt == g->nkt && gc.host.size() == h.size() &&
                    !memcmp(gc.host.data(), h.data(), sizeof(double) * h.size()) && gc.kw == kwv;
  if (!same) {
    if (!gc.num_sm) {
      cudaDeviceProp p;
      CKK(cudaGetDeviceProperties(&p, dev));
      gc.num_sm = p.multiProcessorCount;
    }
    if (gc.valid) CKK(cudaDeviceSynchronize());
    if (gc.cap < h.size()) {
      if (gc.d) cudaFree(gc.d);
      gc.d = nullptr;
      CKK(cudaMalloc(&gc.d, sizeof(double) * h.size()));
      gc.cap = h.size();
    }
    if (!gc.d_kw) CKK(cudaMalloc(&gc.d_kw, sizeof(int) * MAXK));
    gc.host.swap(h);
    gc.kw.swap(kwv);
    gc.nka = g->nka;
    gc.nkt = g->nkt;
    {
      // device copy of qabs: per band the bin-width-weighted pair of str.f90:5144-5145,
      // qabs(ib,jt)*dew(jt) + qabs(ib,jtp)*dew(jtp) (a pure function of the grid, same
      // arithmetic as the reference), laid out [kr][ia][ib][jt] so that a warp's loads of one
      // band are contiguous
      std::vector<double> up(gc.host);
      const size_t q0 = 3 * nka + 3 * nkt + nka * nkt;
      const double *dewh = gc.host.data() + 3 * nka + 2 * nkt;
      for (size_t ra = 0; ra < (size_t)MISTRA_JPTAERRAD * nka; ++ra)
        for (size_t jt = 0; jt < nkt; ++jt) {
          const size_t jtp = jt + 1 < nkt ? jt + 1 : nkt - 1;
          for (size_t ib = 0; ib < (size_t)MB; ++ib) {
            const double qa = gc.host[q0 + (ra * nkt + jt) * MB + ib], qb = gc.host[q0 + (ra * nkt + jtp) * MB + ib];
            const double t0 = qa * dewh[jt], t1 = qb * dewh[jtp];
            up[q0 + (ra * MB + ib) * nkt + jt] = t0 + t1;
          }
        }
      CKK(cudaMemcpyAsync(gc.d, up.data(), sizeof(double) * up.size(), cudaMemcpyHostToDevice, st));
      CKK(cudaStreamSynchronize(st));
    }
    CKK(cudaMemcpyAsync(gc.d_kw, gc.kw.data(), sizeof(int) * nka, cudaMemcpyHostToDevice, st));
    CKK(cudaStreamSynchronize(st));
    gc.valid = true;
  }
  if (!gc.attr_set) {
    CKK(cudaFuncSetAttribute(kon_subkon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    gc.attr_set = true;
  }
  const double *p = gc.d;
  out->nka = g->nka; out->nkt = g->nkt; out->a0m = g->a0m; out->dlne = g->dlne;
  out->en = p; p += nka; out->rn = p; p += nka; out->b0m = p; p += nka;
  out->ew = p; p += nkt; out->e = p; p += nkt; out->dew = p; p += nkt;
  out->rw = p; p += nka * nkt; out->qabs = p; p += (size_t)MISTRA_JPTAERRAD * nka * nkt * MB;
  out->rq = p;
  out->kw = gc.d_kw;
  out->ka = g->ka;
  *cache = &gc;
  return 0;
}

int launch(const mistra_kon_grid *g, int64_t ncell, double dt, const KonArgs &A, void *stream)
{
  const int rch = rows_per_chunk(g);
  if (rch < 1) return mistra_internal_fail(MISTRA_KPP_EINVAL, "grid too large for shared memory");
  const size_t smem = smem_bytes(g) + (((size_t)rch * ((size_t)g->nkt * 10 + 8) + 31) & ~(size_t)15);
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaStream_t st = (cudaStream_t)stream;
  KonGridDev gd;
  GridCache *gc;
  int rc;
  if ((rc = grid_to_device(g, st, &gd, &gc, A.full != 0))) return rc;
  int per_sm = (int)((227 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  long long blocks = (long long)gc->num_sm * per_sm;
  if (blocks > ncell) blocks = ncell;
  kon_subkon_kernel<<<(int)blocks, KON_THREADS, smem, st>>>(gd, ncell, dt, A, rch);
  CKK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // namespace

extern "C" {

int mistra_kon_subkon_device(const mistra_kon_grid *g, int64_t ncell, double dt, double *d_ffk,
                             const double *d_totr, const double *d_dfdt, const double *d_feualt,
                             const double *d_pp, double *d_to, const double *d_tn,
                             double *d_xm1o, const double *d_xm1n, const int32_t *d_kr,
                             int32_t *d_status, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!(dt > 0.0)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "dt <= 0");
  if (ncell == 0) return 0;
  if (!d_ffk || !d_totr || !d_dfdt || !d_feualt || !d_pp || !d_to || !d_tn || !d_xm1o || !d_xm1n || !d_kr)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  KonArgs A;
  memset(&A, 0, sizeof A);
  A.ff = d_ffk; A.totr = d_totr; A.pp = d_pp; A.kr = d_kr; A.status = d_status;
  A.dfdt = d_dfdt; A.feualt = d_feualt; A.tn = d_tn; A.xm1n = d_xm1n; A.to = d_to; A.xm1o = d_xm1o;
  return launch(g, ncell, dt, A, stream);
}

int mistra_kon_subkon(const mistra_kon_grid *g, int64_t ncell, double dt, double *ffk,
                      const double *totr, const double *dfdt, const double *feualt,
                      const double *pp, double *to, const double *tn, double *xm1o,
                      const double *xm1n, const int32_t *kr, int32_t *status, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell == 0) return 0;
  if (!ffk || !totr || !dfdt || !feualt || !pp || !to || !tn || !xm1o || !xm1n || !kr)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  int dev = -1;
  CKK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, b_ff = n * g->nka * g->nkt * 8, b_tr = n * MB * 8, b_s = n * 8, b_i = n * 4;
  const size_t total = b_ff + b_tr + 7 * b_s + 2 * b_i + 64;
  // the per-device scratch buffer is grown, filled and read back under the lock: two host threads on the same
  // device must not free it under each other's copies
  std::lock_guard<std::recursive_mutex> lk_scratch(g_mu);
  Scratch &sc = g_scratch[dev];
  if (sc.bytes < total) {
    if (sc.p) { CKK(cudaDeviceSynchronize()); cudaFree(sc.p); sc.p = nullptr; sc.bytes = 0; }
    CKK(cudaMalloc(&sc.p, total));
    sc.bytes = total;
  }
  char *p = sc.p;
  double *d_ff = (double *)p; p += b_ff;
  double *d_tr = (double *)p; p += b_tr;
  double *d_s[7];
  for (auto &q : d_s) { q = (double *)p; p += b_s; }
  int32_t *d_kr = (int32_t *)p; p += b_i;
  int32_t *d_st = (int32_t *)p;
  const double *hs[7] = {dfdt, feualt, pp, to, tn, xm1o, xm1n};
  CKK(cudaMemcpyAsync(d_ff, ffk, b_ff, cudaMemcpyHostToDevice, st));
  CKK(cudaMemcpyAsync(d_tr, totr, b_tr, cudaMemcpyHostToDevice, st));
  for (int i = 0; i < 7; ++i) CKK(cudaMemcpyAsync(d_s[i], hs[i], b_s, cudaMemcpyHostToDevice, st));
  CKK(cudaMemcpyAsync(d_kr, kr, b_i, cudaMemcpyHostToDevice, st));
  if ((rc = mistra_kon_subkon_device(g, ncell, dt, d_ff, d_tr, d_s[0], d_s[1], d_s[2], d_s[3], d_s[4], d_s[5],
                                     d_s[6], d_kr, status ? d_st : nullptr, stream)))
    return rc;
  CKK(cudaMemcpyAsync(ffk, d_ff, b_ff, cudaMemcpyDeviceToHost, st));
  CKK(cudaMemcpyAsync(to, d_s[3], b_s, cudaMemcpyDeviceToHost, st));
  CKK(cudaMemcpyAsync(xm1o, d_s[5], b_s, cudaMemcpyDeviceToHost, st));
  if (status) CKK(cudaMemcpyAsync(status, d_st, b_i, cudaMemcpyDeviceToHost, st));
  CKK(cudaStreamSynchronize(st));
  return 0;
}

int mistra_kon_layers_device(const mistra_kon_grid *g, int64_t ncell, double dt, int chem,
                             const mistra_kon_state *s, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!(dt > 0.0)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "dt <= 0");
  if (ncell == 0) return 0;
  if (!s || !s->ff || !s->t || !s->talt || !s->xm1 || !s->xm1a || !s->feu || !s->dfddt || !s->xm2 || !s->dtcon ||
      !s->p || !s->totrad || !s->nar)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  if (chem && (!s->vol1_a || !s->vol1_d || !s->part_o_a || !s->part_o_d || !s->part_n_a || !s->part_n_d ||
               !s->vol2 || !s->pntot))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "chem: null sum array");
  KonArgs A;
  memset(&A, 0, sizeof A);
  A.full = 1; A.chem = chem ? 1 : 0;
  A.ff = s->ff; A.totr = s->totrad; A.pp = s->p; A.kr = s->nar; A.status = s->status;
  A.t = s->t; A.talt = s->talt; A.xm1 = s->xm1; A.xm1a = s->xm1a; A.feu = s->feu; A.dfddt = s->dfddt;
  A.xm2 = s->xm2; A.dtcon = s->dtcon;
  A.vol1_a = s->vol1_a; A.vol1_d = s->vol1_d; A.part_o_a = s->part_o_a; A.part_o_d = s->part_o_d;
  A.part_n_a = s->part_n_a; A.part_n_d = s->part_n_d; A.vol2 = s->vol2; A.pntot = s->pntot;
  return launch(g, ncell, dt, A, stream);
}

int mistra_kon_layers(const mistra_kon_grid *g, int64_t ncell, double dt, int chem,
                      const mistra_kon_state *s, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell == 0) return 0;
  if (!s) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null state");
  int dev = -1;
  CKK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, nka = g->nka, b_ff = n * nka * g->nkt * 8, b_s = n * 8;
  // staging plan: (host pointer, bytes, copy in, copy out)
  struct Item { void *h; size_t bytes; bool in, out; void **slot; };
  mistra_kon_state d = *s;
  std::vector<Item> items = {
      {s->ff, b_ff, true, true, (void **)&d.ff}, {s->t, b_s, true, true, (void **)&d.t},
      {s->talt, b_s, true, true, (void **)&d.talt}, {s->xm1, b_s, true, true, (void **)&d.xm1},
      {s->xm1a, b_s, true, true, (void **)&d.xm1a}, {s->feu, b_s, true, true, (void **)&d.feu},
      {s->dfddt, b_s, true, true, (void **)&d.dfddt}, {s->xm2, b_s, true, true, (void **)&d.xm2},
      {s->dtcon, b_s, true, true, (void **)&d.dtcon}, {(void *)s->p, b_s, true, false, (void **)&d.p},
      {(void *)s->totrad, n * MB * 8, true, false, (void **)&d.totrad},
      {(void *)s->nar, n * 4, true, false, (void **)&d.nar},
      {s->status, n * 4, false, true, (void **)&d.status}};
  if (chem) {
    void **sl[6] = {(void **)&d.vol1_a, (void **)&d.vol1_d, (void **)&d.part_o_a, (void **)&d.part_o_d,
                    (void **)&d.part_n_a, (void **)&d.part_n_d};
    double *hp[6] = {s->vol1_a, s->vol1_d, s->part_o_a, s->part_o_d, s->part_n_a, s->part_n_d};
    for (int i = 0; i < 6; ++i) items.push_back({hp[i], n * nka * 8, false, true, sl[i]});
    items.push_back({s->vol2, n * 4 * 8, false, true, (void **)&d.vol2});
    items.push_back({s->pntot, n * 4 * 8, false, true, (void **)&d.pntot});
  }
  size_t total = 0;
  for (auto &it : items)
    if (it.h) total += (it.bytes + 255) & ~(size_t)255;
  // the per-device scratch buffer is grown, filled and read back under the lock: two host threads on the same
  // device must not free it under each other's copies
  std::lock_guard<std::recursive_mutex> lk_scratch(g_mu);
  Scratch &sc = g_scratch[dev];
  if (sc.bytes < total) {
    if (sc.p) { CKK(cudaDeviceSynchronize()); cudaFree(sc.p); sc.p = nullptr; sc.bytes = 0; }
    CKK(cudaMalloc(&sc.p, total));
    sc.bytes = total;
  }
  char *p = sc.p;
  for (auto &it : items) {
    if (!it.h) { *it.slot = nullptr; continue; }
    *it.slot = p;
    if (it.in) CKK(cudaMemcpyAsync(p, it.h, it.bytes, cudaMemcpyHostToDevice, st));
    p += (it.bytes + 255) & ~(size_t)255;
  }
  if ((rc = mistra_kon_layers_device(g, ncell, dt, chem, &d, stream))) return rc;
  for (auto &it : items)
    if (it.h && it.out) CKK(cudaMemcpyAsync(it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKK(cudaStreamSynchronize(st));
  return 0;
}

int64_t mistra_kon_launch_count(void) { return g_launches.load(); }

#ifdef KON_PROF
int mistra_kon_prof(unsigned long long *out)
{
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, g_kon_prof, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(g_kon_prof, z, sizeof z);
  return (int)e;
}
#endif

}  // extern "C"
