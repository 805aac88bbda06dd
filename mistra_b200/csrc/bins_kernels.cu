// 2-D aerosol/droplet bin redistribution around the chemistry step (include/mistra_bins.h):
// CUDA kernels + C-ABI entries.  Role in the reference: the loop nests of SUBROUTINE
// stem_kpp, /root/reference/src/str.f90:5916-5966 (snapshot) and 5976-6134
// (redistribution + exchange of dissolved species).
//
// Mapping: one CTA per model layer ("cell"), the layer's ff(jt,ia) tile (nka x nkt
// doubles, 39.2 kB for 70 x 70) staged in shared memory by coalesced loads, written back
// once.  The kernels are HBM-bound: 2 x 39.2 kB (+ 11 kB of sl1/sion1) per layer-call.
//   * Sums whose rounding feeds the result (sap, smp -> x0 -> ix/c0 -> ff) are taken in the
//     reference's loop order by one thread per chem bin, with explicit __dadd_rn/__dmul_rn
//     (no FMA contraction), so ff is reproduced to the last bit.
//   * The shift along the dry-mass axis is sequential in ia (in-place, direction depends
//     on the sign of the mass change, str.f90:6012-6021) but independent between water
//     bins jt: one thread per jt walks ia in the reference's order.
//   * Transferred volumes vc(tix,kc) are accumulated per jt and reduced in jt order.
#include "../../include/mistra_bins.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int NKC = MISTRA_NKC, LSP = MISTRA_LSP, J2 = MISTRA_J2, J6 = MISTRA_J6;
constexpr int BINS_THREADS = 128;
constexpr int MAXK = 128;  // max nka / nkt

__constant__ int c_lj2[LSP] = {1, 2, 8, 9, 13, 14, 19, 20, 30};  // str.f90:5876

struct GridDev {
  int nka, nkt, ka, nkc_l, ial_first;
  const int *kw;
  const double *en;
  const double *rq;
};

__device__ __forceinline__ void ia_range(const GridDev &g, int kc, int &ial, int &iau)
{
  if (kc == 1 || kc == 3) { ial = g.ial_first; iau = g.ka; }  // str.f90:5924-5931
  else { ial = g.ka + 1; iau = g.nka - 1; }                   // str.f90:5932-5935
}

// Asynchronous global->shared tile copies (cp.async, LDGSTS in SASS): no registers, and
// the copy overlaps the per-layer scalar work issued before tile_wait().
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
__device__ __forceinline__ void tile_wait()
{
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void tile_load_async(double *__restrict__ dst, const double *__restrict__ src, int n)
{
  if ((n & 1) == 0)
    for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) cp_async16(dst + 2 * i, src + 2 * i);
  else
    for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async8(dst + i, src + i);
}
__device__ __forceinline__ void tile_load_padded_async(double *__restrict__ dst, const double *__restrict__ src,
                                                       int n, int nkt, int TS)
{
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r0 = i / nkt;
    cp_async8(dst + r0 * TS + (i - r0 * nkt), src + i);
  }
}

// Coalesced 16-byte write-back of a tile.
__device__ __forceinline__ void tile_store(double *__restrict__ dst, const double *__restrict__ src, int n)
{
  if ((n & 1) == 0) {
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) d2[i] = s2[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  }
}

// ---- str.f90:5916-5966 ---------------------------------------------------------------
__global__ void __launch_bounds__(BINS_THREADS)
bins_snapshot_kernel(GridDev g, long long ncell, const double *__restrict__ ff,
                     const double *__restrict__ cm, const double *__restrict__ sion1,
                     double *__restrict__ sap, double *__restrict__ smp,
                     double *__restrict__ sion1o)
{
  extern __shared__ __align__(16) double sm[];
  const int nka = g.nka, nkt = g.nkt, ntile = nka * nkt, TS = nkt | 1;
  double *t_ff = sm;                         // [nka][TS]
  double *s_en = sm + nka * TS;
  int *s_kw = (int *)(s_en + nka);
  for (int i = threadIdx.x; i < nka; i += blockDim.x) { s_en[i] = g.en[i]; s_kw[i] = g.kw[i]; }
  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    bool any = false;
    for (int kc = 0; kc < g.nkc_l; ++kc) any |= (cm[c * NKC + kc] != 0.0);
    __syncthreads();  // previous tile fully consumed
    if (any) {
      tile_load_padded_async(t_ff, ff + (size_t)c * ntile, ntile, nkt, TS);
      tile_wait();
    }
    __syncthreads();
    // fs(ia,kc) = sum_jt ff*en is an independent chain per dry class: one thread per
    // (kc parity, ia); the long chains (sap over all (ia,jt) of the bin, smp over ia) stay
    // sequential in the reference's order on one thread per chem bin.
    double *s_fs = (double *)(s_kw + nka + (nka & 1));   // [2][nka]: aerosol part / droplet part
    double *s_ns = s_fs + 2 * nka;                       // [2][nka]: particle number per class
    for (int p = threadIdx.x; p < 2 * nka; p += blockDim.x) {
      const int part = p / nka, ia = p % nka + 1;
      int jtl, jtu;
      if (part == 0) { jtl = 1; jtu = s_kw[ia - 1]; }
      else { jtl = s_kw[ia - 1] + 1; jtu = nkt; }
      const double en = s_en[ia - 1];
      const double *row = t_ff + (ia - 1) * TS;
      double fs = 0.0, ns = 0.0;
      for (int jt = jtl; jt <= jtu; ++jt) {
        const double v = row[jt - 1];
        fs = __dadd_rn(fs, __dmul_rn(v, en));        // str.f90:5948
        ns = __dadd_rn(ns, v);
      }
      s_fs[p] = fs;
      s_ns[p] = ns;
    }
    __syncthreads();
    const int kc = threadIdx.x + 1;
    if (kc <= NKC) {
      double sapk = 0.0, smpk = 0.0;
      if (kc <= g.nkc_l && cm[c * NKC + kc - 1] != 0.0) {
        int ial, iau;
        ia_range(g, kc, ial, iau);
        for (int ia = ial; ia <= iau; ++ia) {
#ifdef KPP_STRICT
          // the reference's running sum over (ia, jt) (str.f90:5949), one rounding per particle bin
          int jtl, jtu;
          if (kc <= 2) { jtl = 1; jtu = s_kw[ia - 1]; }
          else { jtl = s_kw[ia - 1] + 1; jtu = nkt; }
          const double *row = t_ff + (ia - 1) * TS;
          for (int jt = jtl; jt <= jtu; ++jt) sapk = __dadd_rn(sapk, row[jt - 1]);
#else
          // per-class partial sums first: the 1600-long dependent chain of FP64 adds
          // (~45 cycles each) becomes 70 chains of <= 70 in parallel + one of <= 70
          sapk = __dadd_rn(sapk, s_ns[(kc <= 2 ? 0 : nka) + ia - 1]);
#endif
          smpk = __dadd_rn(smpk, s_fs[(kc <= 2 ? 0 : nka) + ia - 1]);   // str.f90:5951
        }
        for (int l = 0; l < LSP; ++l)              // str.f90:5959-5962
          sion1o[(c * NKC + kc - 1) * LSP + l] = sion1[(c * NKC + kc - 1) * J6 + c_lj2[l] - 1];
      }
      sap[c * NKC + kc - 1] = sapk;
      smp[c * NKC + kc - 1] = smpk;
    }
  }
}

// ---- str.f90:5976-6134 ---------------------------------------------------------------
__global__ void __launch_bounds__(BINS_THREADS)
bins_redistribute_kernel(GridDev g, long long ncell, double *__restrict__ ff,
                         const double *__restrict__ cm, const double *__restrict__ cw,
                         const double *__restrict__ sap, const double *__restrict__ smp,
                         const double *__restrict__ sion1o, double *__restrict__ sion1,
                         double *__restrict__ sl1, int *__restrict__ nwarn)
{
  extern __shared__ __align__(16) double sm[];
  const int nka = g.nka, nkt = g.nkt, ntile = nka * nkt;
  double *t_ff = sm;                         // [nka][nkt]
  double *s_en = t_ff + ntile;               // [nka]
  double *s_c0 = s_en + nka;                 // [NKC][nka]
  double *s_vcw = s_c0 + NKC * nka;          // [warps][NKC*NKC] per-warp partial volumes
  double *s_vc = s_vcw + (BINS_THREADS / 32) * NKC * NKC;   // [NKC*NKC]
  double *s_den = s_vc + NKC * NKC;          // [NKC]
  double *s_xf = s_den + NKC;                // [NKC*NKC] exchange fraction per (from, to)
  double *s_ds = s_xf + NKC * NKC;           // [NKC][LSP]
  double *s_c1 = s_ds + NKC * LSP;           // [NKC][nka]  1 - c0
  int *s_kw = (int *)(s_c1 + NKC * nka);     // [nka]
  int *s_ix = s_kw + nka;                    // [NKC][nka]
  int *s_act = s_ix + NKC * nka;             // [NKC] + warn counter
  const double fpi = 4.0 / 3.0 * 3.1415926535897932;  // str.f90:5838
  const double em6 = (double)1.e-06f;                  // default-REAL literal of str.f90:5983
  for (int i = threadIdx.x; i < nka; i += blockDim.x) { s_en[i] = g.en[i]; s_kw[i] = g.kw[i]; }
  // bit ia-1 of m_aer: water bin jt = threadIdx.x + 1 of dry class ia is "aerosol" (jt <= kw(ia));
  // the walk of phase 2 then visits only the classes whose jt belongs to the chem bin
  unsigned m_aer[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {                    // static indexing keeps the masks in registers
    unsigned m = 0u;
    for (int b = 0; b < 32; ++b) {
      const int ia = 32 * q + b;
      if (ia < nka && (int)threadIdx.x + 1 <= g.kw[ia]) m |= 1u << b;
    }
    m_aer[q] = m;
  }

  for (long long c = blockIdx.x; c < ncell; c += gridDim.x) {
    __syncthreads();
    bool any = false;
    for (int kc = 0; kc < g.nkc_l; ++kc) any |= (cm[c * NKC + kc] != 0.0 && sap[c * NKC + kc] > 1.e-6);
    double *f = ff + (size_t)c * ntile;
    if (any) tile_load_async(t_ff, f, ntile);    // in flight during phases 0 and 1
    // ---- phase 0: mass change per particle of every bin (str.f90:5979-5993) ----
    if (threadIdx.x < NKC * LSP) {           // the 36 ion differences in parallel
      const int kc = threadIdx.x / LSP + 1, l = threadIdx.x % LSP;
      double v = 0.0;
      if (kc <= g.nkc_l && cm[c * NKC + kc - 1] != 0.0 && sap[c * NKC + kc - 1] > 1.e-6) {
        const double d = __dadd_rn(sion1[(c * NKC + kc - 1) * J6 + c_lj2[l] - 1],
                                   -sion1o[(c * NKC + kc - 1) * LSP + l]);
        v = __ddiv_rn(__dmul_rn(d, em6), sap[c * NKC + kc - 1]);
      }
      s_ds[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < NKC) {
      const int kc = threadIdx.x + 1;
      int act = 0;
      double den = 0.0;
      if (kc <= g.nkc_l && cm[c * NKC + kc - 1] != 0.0 && sap[c * NKC + kc - 1] > 1.e-6) {
        act = 1;
        const double mw[LSP] = {1., 18., 96., 44., 62., 35.5, 97., 23., 95.};
        const double *ds = s_ds + (kc - 1) * LSP;
        double sacc = __dmul_rn(ds[0], mw[0]);
        for (int l = 1; l < LSP; ++l) sacc = __dadd_rn(sacc, __dmul_rn(ds[l], mw[l]));
        den = __dmul_rn(sacc, 1000.);
      }
      s_act[threadIdx.x] = act;
      s_den[threadIdx.x] = den;
    }
    if (threadIdx.x == NKC) { s_act[NKC] = 0; s_act[NKC + 1] = 0; }  // warn counter, transfer flag
    for (int i = threadIdx.x; i < (BINS_THREADS / 32) * NKC * NKC; i += blockDim.x) s_vcw[i] = 0.0;
    __syncthreads();
    if (any) {
      // ---- phase 1: target class ix and split c0 of every (kc, ia) (str.f90:6023-6044) ----
      for (int p = threadIdx.x; p < NKC * nka; p += blockDim.x) {
        const int kc = p / nka + 1, ia = p % nka + 1;
        int ix = 0;
        double c0 = 0.0;
        if (s_act[kc - 1]) {
          int ial, iau;
          ia_range(g, kc, ial, iau);
          if (ia >= ial && ia <= iau) {
            const double den = s_den[kc - 1], en = s_en[ia - 1];
            const double x0 = __dadd_rn(en, __dmul_rn(__ddiv_rn(__dmul_rn(den, en), smp[c * NKC + kc - 1]),
                                                      sap[c * NKC + kc - 1]));
            if (!(den > 0.0) && x0 <= 0.0) atomicAdd(&s_act[NKC], 1);
            // en is strictly increasing, so the reference's linear scan for
            // en(iia) <= x0 < en(iia+1) has at most one hit: bisect for it
            if (x0 >= s_en[0] && x0 < s_en[nka - 1]) {
              int lo = 1, hi = nka;                  // invariant: en(lo) <= x0 < en(hi)
              while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_en[mid - 1] <= x0) lo = mid; else hi = mid;
              }
              ix = lo;
              c0 = __ddiv_rn(__dadd_rn(s_en[lo], -x0), __dadd_rn(s_en[lo], -s_en[lo - 1]));
            }
            if (ix == 0) {
              if (s_en[0] > x0) { ix = 1; c0 = 1.0; }
              else { ix = nka - 1; c0 = 0.0; }
            }
          }
        }
        s_ix[p] = ix;
        s_c0[p] = c0;
        s_c1[p] = __dadd_rn(1.0, -c0);
      }
      tile_wait();
      __syncthreads();
      // ---- phase 2: one thread per water bin jt walks the dry classes (str.f90:6012-6096) ----
      const int jt = threadIdx.x + 1;
      {
        for (int kc = 1; kc <= g.nkc_l; ++kc) {
          if (!s_act[kc - 1]) continue;
          double vq1 = 0.0, vq2 = 0.0, vq3 = 0.0, vq4 = 0.0;   // volume moved to chem bins 1..4 by this jt
          int ial, iau;
          ia_range(g, kc, ial, iau);
          const int iinkr = (s_den[kc - 1] >= 0.0) ? -1 : 1;          // str.f90:6016-6020
          // classes ial..iau whose water bin jt belongs to this chem bin, as bit masks
          unsigned mk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int lo = ial - 1 - 32 * q, hi = iau - 32 * q;          // bits [lo, hi) of word q
            unsigned range = (hi <= 0 || lo >= 32) ? 0u
                             : ((hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & (lo <= 0 ? 0xffffffffu : ~((1u << lo) - 1u)));
            mk[q] = (jt <= nkt) ? (range & (kc <= 2 ? m_aer[q] : ~m_aer[q])) : 0u;
          }
          for (;;) {
            int ia;                                                       // next class in walk order
            if (iinkr > 0) {
              if (mk[0]) { const int b = __ffs(mk[0]) - 1; mk[0] &= mk[0] - 1; ia = b + 1; }
              else if (mk[1]) { const int b = __ffs(mk[1]) - 1; mk[1] &= mk[1] - 1; ia = b + 33; }
              else if (mk[2]) { const int b = __ffs(mk[2]) - 1; mk[2] &= mk[2] - 1; ia = b + 65; }
              else if (mk[3]) { const int b = __ffs(mk[3]) - 1; mk[3] &= mk[3] - 1; ia = b + 97; }
              else break;
            } else {
              if (mk[3]) { const int b = 31 - __clz(mk[3]); mk[3] ^= 1u << b; ia = b + 97; }
              else if (mk[2]) { const int b = 31 - __clz(mk[2]); mk[2] ^= 1u << b; ia = b + 65; }
              else if (mk[1]) { const int b = 31 - __clz(mk[1]); mk[1] ^= 1u << b; ia = b + 33; }
              else if (mk[0]) { const int b = 31 - __clz(mk[0]); mk[0] ^= 1u << b; ia = b + 1; }
              else break;
            }
            const double x1 = t_ff[(ia - 1) * nkt + jt - 1];
            // x1 > 0 on the integer pipe (an FP64 compare costs ~45 cycles here and most bins are
            // empty): positive doubles are exactly the positive int64 patterns, NaNs excluded
            const long long xb = __double_as_longlong(x1);
            if (xb > 0 && xb <= 0x7ff0000000000000LL) {
              const int ix = s_ix[(kc - 1) * nka + ia - 1];
              const double c0 = s_c0[(kc - 1) * nka + ia - 1];
              const double a = __dmul_rn(x1, c0), b = __dmul_rn(x1, s_c1[(kc - 1) * nka + ia - 1]);
              t_ff[(ia - 1) * nkt + jt - 1] = 0.0;
              t_ff[(ix - 1) * nkt + jt - 1] = __dadd_rn(t_ff[(ix - 1) * nkt + jt - 1], a);
              t_ff[ix * nkt + jt - 1] = __dadd_rn(t_ff[ix * nkt + jt - 1], b);
              int tix, tixp;
              if (ix > g.ka) tix = (jt > s_kw[ix - 1]) ? 4 : 2;
              else tix = (jt > s_kw[ix - 1]) ? 3 : 1;
              if (ix + 1 > g.ka) tixp = (jt > s_kw[ix]) ? 4 : 2;
              else tixp = (jt > s_kw[ix]) ? 3 : 1;
              double r3 = 0.0;
              if (tix != kc || tixp != kc) {        // rare: the move crosses a chem-bin limit
                const double r = g.rq[(ia - 1) * nkt + jt - 1];
                r3 = __dmul_rn(__dmul_rn(r, r), r);
              }
              if (tix != kc) {
                const double dv = __dmul_rn(__dmul_rn(a, fpi), r3);
                if (tix == 1) vq1 = __dadd_rn(vq1, dv); else if (tix == 2) vq2 = __dadd_rn(vq2, dv);
                else if (tix == 3) vq3 = __dadd_rn(vq3, dv); else vq4 = __dadd_rn(vq4, dv);
              }
              if (tixp != kc) {
                const double dv = __dmul_rn(__dmul_rn(b, fpi), r3);
                if (tixp == 1) vq1 = __dadd_rn(vq1, dv); else if (tixp == 2) vq2 = __dadd_rn(vq2, dv);
                else if (tixp == 3) vq3 = __dadd_rn(vq3, dv); else vq4 = __dadd_rn(vq4, dv);
              }
            }
          }
          // fixed-tree reduction over the warp's water bins (deterministic), one partial per warp
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            vq1 = __dadd_rn(vq1, __shfl_xor_sync(0xffffffffu, vq1, off));
            vq2 = __dadd_rn(vq2, __shfl_xor_sync(0xffffffffu, vq2, off));
            vq3 = __dadd_rn(vq3, __shfl_xor_sync(0xffffffffu, vq3, off));
            vq4 = __dadd_rn(vq4, __shfl_xor_sync(0xffffffffu, vq4, off));
          }
          if ((threadIdx.x & 31) == 0) {
            double *o = s_vcw + (threadIdx.x >> 5) * NKC * NKC + (kc - 1) * NKC;
            o[0] = vq1; o[1] = vq2; o[2] = vq3; o[3] = vq4;
          }
        }
      }
      __syncthreads();
      // ---- phase 3: transferred volume per (from, to) pair: the warps' partials in warp order ----
      if (threadIdx.x < NKC * NKC) {
        double s = 0.0;
        for (int wq = 0; wq < BINS_THREADS / 32; ++wq) s = __dadd_rn(s, s_vcw[wq * NKC * NKC + threadIdx.x]);
        s_vc[threadIdx.x] = s;
      }
      tile_store(f, t_ff, ntile);
      __syncthreads();
      // ---- phase 4: exchange of dissolved species between bins (str.f90:6102-6134) ----
      // exchange fraction of every (from, to) pair once (0 = no transfer) ...
      if (threadIdx.x < NKC * NKC) {
        const int kc = threadIdx.x / NKC, kkc = threadIdx.x % NKC;
        double xfact = 0.0;
        const double vcv = s_vc[threadIdx.x];
        if (kc != kkc && kc < g.nkc_l && kkc < g.nkc_l && vcv != 0.0) {
          const double cwf = cw[c * NKC + kc];
          if (cwf > 0.0) {
            const double vol_ch = __dmul_rn(vcv, 1.e-12);
            xfact = __dadd_rn(1.0, -__ddiv_rn(__dadd_rn(cwf, -vol_ch), cwf));
            s_act[NKC + 1] = 1;                       // at least one transfer in this layer
          }
        }
        s_xf[threadIdx.x] = xfact;
      }
      __syncthreads();
      // ... then every species l of sl1 (121) and sion1 (55) walks the pairs in the reference's order
      if (s_act[NKC + 1]) {
        for (int l = threadIdx.x; l < J2 + J6; l += blockDim.x) {
          const int nl = l < J2 ? J2 : J6, ll = l < J2 ? l : l - J2;
          double *base = (l < J2 ? sl1 : sion1) + (size_t)c * NKC * nl;
          double v[NKC];
#pragma unroll
          for (int k = 0; k < NKC; ++k) v[k] = base[k * nl + ll];
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc)
#pragma unroll
            for (int kkc = 0; kkc < NKC; ++kkc) {            // pairs outside nkc_l have xfact = 0
              const double xfact = s_xf[kc * NKC + kkc];
              if (kkc == kc || xfact == 0.0) continue;     // xfact == 0 moves nothing (xch = 0)
              const double xch = __dmul_rn(v[kc], xfact);
              v[kc] = __dadd_rn(v[kc], -xch);
              v[kkc] = __dadd_rn(v[kkc], xch);
            }
#pragma unroll
          for (int k = 0; k < NKC; ++k) base[k * nl + ll] = v[k];
        }
      }
    }
    if (nwarn && threadIdx.x == 0) nwarn[c] = any ? s_act[NKC] : 0;
  }
}

// ---- host side -------------------------------------------------------------------------
std::mutex g_mu;
std::atomic<long long> g_launches{0};

struct GridCache {
  bool valid = false;
  int nka = 0, nkt = 0;
  std::vector<int> kw;
  std::vector<double> en, rq;
  int *d_kw = nullptr;
  double *d_en = nullptr, *d_rq = nullptr;
  int num_sm = 0;
  bool attr_set = false;
};
GridCache g_cache[16];

#define CKB(call)                                                                      \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return mistra_internal_fail(e_ == cudaErrorMemoryAllocation ? MISTRA_KPP_ENOMEM  \
                                  : (e_ == cudaErrorNoDevice ? MISTRA_KPP_ENODEVICE    \
                                                             : MISTRA_KPP_ECUDA),      \
                                  std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int check_grid(const mistra_bins_grid *g)
{
  if (!g || !g->kw || !g->en || !g->rq) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null grid");
  if (g->nka < 2 || g->nka > MAXK || g->nkt < 1 || g->nkt > BINS_THREADS)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "grid size out of range (nka <= 128, nkt <= 128)");
  if (g->nkc_l < 1 || g->nkc_l > NKC || g->ka < 0 || g->ka > g->nka || g->ial_first < 1 || g->ial_first > 2)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad ka / nkc_l / ial_first");
  for (int i = 0; i < g->nka; ++i)
    if (g->kw[i] < 0 || g->kw[i] > g->nkt) return mistra_internal_fail(MISTRA_KPP_EINVAL, "kw out of range");
  return 0;
}

// upload the grid arrays once per distinct grid (stream-ordered; a changed grid waits for
// the kernels that still read the old one)
int grid_to_device(const mistra_bins_grid *g, cudaStream_t st, GridDev *out, GridCache **cache)
{
  int dev = -1;
  CKB(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  GridCache &gc = g_cache[dev];
  const size_t nrq = (size_t)g->nka * g->nkt;
  const bool same = gc.valid && gc.nka == g->nka && gc.nkt == g->nkt &&
                    !memcmp(gc.kw.data(), g->kw, sizeof(int) * g->nka) &&
                    !memcmp(gc.en.data(), g->en, sizeof(double) * g->nka) &&
                    !memcmp(gc.rq.data(), g->rq, sizeof(double) * nrq);
  if (!same) {
    if (!gc.d_kw) {
      cudaDeviceProp p;
      CKB(cudaGetDeviceProperties(&p, dev));
      gc.num_sm = p.multiProcessorCount;
      CKB(cudaMalloc(&gc.d_kw, sizeof(int) * MAXK));
      CKB(cudaMalloc(&gc.d_en, sizeof(double) * MAXK));
      CKB(cudaMalloc(&gc.d_rq, sizeof(double) * MAXK * MAXK));
    }
    if (gc.valid) CKB(cudaDeviceSynchronize());
    gc.kw.assign(g->kw, g->kw + g->nka);
    gc.en.assign(g->en, g->en + g->nka);
    gc.rq.assign(g->rq, g->rq + nrq);
    gc.nka = g->nka;
    gc.nkt = g->nkt;
    CKB(cudaMemcpyAsync(gc.d_kw, gc.kw.data(), sizeof(int) * g->nka, cudaMemcpyHostToDevice, st));
    CKB(cudaMemcpyAsync(gc.d_en, gc.en.data(), sizeof(double) * g->nka, cudaMemcpyHostToDevice, st));
    CKB(cudaMemcpyAsync(gc.d_rq, gc.rq.data(), sizeof(double) * nrq, cudaMemcpyHostToDevice, st));
    CKB(cudaStreamSynchronize(st));
    gc.valid = true;
  }
  if (!gc.attr_set) {
    CKB(cudaFuncSetAttribute(bins_snapshot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CKB(cudaFuncSetAttribute(bins_redistribute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    gc.attr_set = true;
  }
  out->nka = g->nka; out->nkt = g->nkt; out->ka = g->ka; out->nkc_l = g->nkc_l;
  out->ial_first = g->ial_first;
  out->kw = gc.d_kw; out->en = gc.d_en; out->rq = gc.d_rq;
  *cache = &gc;
  return 0;
}

// device staging of the host-buffer entries, grown on demand and kept
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_scratch[16];

int scratch(size_t bytes, char **out)
{
  int dev = -1;
  CKB(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  Scratch &sc = g_scratch[dev];
  if (sc.bytes < bytes) {
    if (sc.p) { CKB(cudaDeviceSynchronize()); cudaFree(sc.p); sc.p = nullptr; sc.bytes = 0; }
    CKB(cudaMalloc(&sc.p, bytes));
    sc.bytes = bytes;
  }
  *out = sc.p;
  return 0;
}

size_t tile_stride(int nkt) { return (size_t)(nkt | 1); }

size_t smem_snapshot(const mistra_bins_grid *g)
{
  return sizeof(double) * ((size_t)g->nka * tile_stride(g->nkt) + g->nka + 4 * g->nka) + sizeof(int) * (g->nka + 2);
}
size_t smem_redistribute(const mistra_bins_grid *g)
{
  const size_t nka = g->nka, nkt = g->nkt;
  return sizeof(double) * (nka * nkt + nka + 2 * NKC * nka + (BINS_THREADS / 32 + 2) * NKC * NKC + NKC + NKC * LSP) +
         sizeof(int) * (nka + NKC * nka + NKC + 4);
}

int grid_blocks(const GridCache &gc, int64_t ncell, size_t smem)
{
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long b = (long long)gc.num_sm * per_sm;
  return (int)(ncell < b ? ncell : b);
}

}  // namespace

extern "C" {

int mistra_bins_snapshot_device(const mistra_bins_grid *g, int64_t ncell, const double *d_ff,
                                const double *d_cm, const double *d_sion1, double *d_sap,
                                double *d_smp, double *d_sion1o, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell == 0) return 0;
  if (!d_ff || !d_cm || !d_sion1 || !d_sap || !d_smp || !d_sion1o)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  std::lock_guard<std::mutex> lk(g_mu);
  cudaStream_t st = (cudaStream_t)stream;
  GridDev gd;
  GridCache *gc;
  if ((rc = grid_to_device(g, st, &gd, &gc))) return rc;
  const size_t smem = smem_snapshot(g);
  bins_snapshot_kernel<<<grid_blocks(*gc, ncell, smem), BINS_THREADS, smem, st>>>(
      gd, ncell, d_ff, d_cm, d_sion1, d_sap, d_smp, d_sion1o);
  CKB(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int mistra_bins_redistribute_device(const mistra_bins_grid *g, int64_t ncell, double *d_ff,
                                    const double *d_cm, const double *d_cw, const double *d_sap,
                                    const double *d_smp, const double *d_sion1o, double *d_sion1,
                                    double *d_sl1, int32_t *d_nwarn, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell == 0) return 0;
  if (!d_ff || !d_cm || !d_cw || !d_sap || !d_smp || !d_sion1o || !d_sion1 || !d_sl1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  std::lock_guard<std::mutex> lk(g_mu);
  cudaStream_t st = (cudaStream_t)stream;
  GridDev gd;
  GridCache *gc;
  if ((rc = grid_to_device(g, st, &gd, &gc))) return rc;
  const size_t smem = smem_redistribute(g);
  bins_redistribute_kernel<<<grid_blocks(*gc, ncell, smem), BINS_THREADS, smem, st>>>(
      gd, ncell, d_ff, d_cm, d_cw, d_sap, d_smp, d_sion1o, d_sion1, d_sl1, d_nwarn);
  CKB(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

// Host-buffer entries: stage, run, copy back (synchronous).
int mistra_bins_snapshot(const mistra_bins_grid *g, int64_t ncell, const double *ff,
                         const double *cm, const double *sion1, double *sap, double *smp,
                         double *sion1o, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell == 0) return 0;
  if (!ff || !cm || !sion1 || !sap || !smp || !sion1o) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, tile = (size_t)g->nka * g->nkt;
  const size_t b_ff = n * tile * 8, b_c = n * NKC * 8, b_si = n * NKC * J6 * 8, b_so = n * NKC * LSP * 8;
  char *d = nullptr;
  if ((rc = scratch(b_ff + 3 * b_c + b_si + b_so, &d))) return rc;
  double *d_ff = (double *)d, *d_cm = (double *)(d + b_ff), *d_sap = d_cm + n * NKC, *d_smp = d_sap + n * NKC;
  double *d_si = d_smp + n * NKC, *d_so = (double *)((char *)d_si + b_si);
  auto done = [&](int r) { return r; };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(mistra_internal_fail(MISTRA_KPP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
  CKF(cudaMemcpyAsync(d_ff, ff, b_ff, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_cm, cm, b_c, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_si, sion1, b_si, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_so, sion1o, b_so, cudaMemcpyHostToDevice, st));  // entries of inactive bins stay as given
  if ((rc = mistra_bins_snapshot_device(g, ncell, d_ff, d_cm, d_si, d_sap, d_smp, d_so, stream))) return done(rc);
  CKF(cudaMemcpyAsync(sap, d_sap, b_c, cudaMemcpyDeviceToHost, st));
  CKF(cudaMemcpyAsync(smp, d_smp, b_c, cudaMemcpyDeviceToHost, st));
  CKF(cudaMemcpyAsync(sion1o, d_so, b_so, cudaMemcpyDeviceToHost, st));
  CKF(cudaStreamSynchronize(st));
  return done(0);
}

int mistra_bins_redistribute(const mistra_bins_grid *g, int64_t ncell, double *ff,
                             const double *cm, const double *cw, const double *sap,
                             const double *smp, const double *sion1o, double *sion1,
                             double *sl1, int32_t *nwarn, void *stream)
{
  int rc = check_grid(g);
  if (rc) return rc;
  if (ncell < 0) return mistra_internal_fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell == 0) return 0;
  if (!ff || !cm || !cw || !sap || !smp || !sion1o || !sion1 || !sl1)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)ncell, tile = (size_t)g->nka * g->nkt;
  const size_t b_ff = n * tile * 8, b_c = n * NKC * 8, b_si = n * NKC * J6 * 8, b_so = n * NKC * LSP * 8,
               b_sl = n * NKC * J2 * 8, b_w = n * 4;
  char *d = nullptr;
  if ((rc = scratch(b_ff + 4 * b_c + b_si + b_so + b_sl + b_w, &d))) return rc;
  double *d_ff = (double *)d, *d_cm = (double *)(d + b_ff), *d_cw = d_cm + n * NKC, *d_sap = d_cw + n * NKC,
         *d_smp = d_sap + n * NKC;
  double *d_si = d_smp + n * NKC, *d_so = (double *)((char *)d_si + b_si), *d_sl = (double *)((char *)d_so + b_so);
  int32_t *d_w = (int32_t *)((char *)d_sl + b_sl);
  auto done = [&](int r) { return r; };
  CKF(cudaMemcpyAsync(d_ff, ff, b_ff, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_cm, cm, b_c, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_cw, cw, b_c, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_sap, sap, b_c, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_smp, smp, b_c, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_si, sion1, b_si, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_so, sion1o, b_so, cudaMemcpyHostToDevice, st));
  CKF(cudaMemcpyAsync(d_sl, sl1, b_sl, cudaMemcpyHostToDevice, st));
  if ((rc = mistra_bins_redistribute_device(g, ncell, d_ff, d_cm, d_cw, d_sap, d_smp, d_so, d_si, d_sl,
                                            nwarn ? d_w : nullptr, stream)))
    return done(rc);
  CKF(cudaMemcpyAsync(ff, d_ff, b_ff, cudaMemcpyDeviceToHost, st));
  CKF(cudaMemcpyAsync(sion1, d_si, b_si, cudaMemcpyDeviceToHost, st));
  CKF(cudaMemcpyAsync(sl1, d_sl, b_sl, cudaMemcpyDeviceToHost, st));
  if (nwarn) CKF(cudaMemcpyAsync(nwarn, d_w, b_w, cudaMemcpyDeviceToHost, st));
  CKF(cudaStreamSynchronize(st));
  return done(0);
#undef CKF
}

int64_t mistra_bins_launch_count(void) { return g_launches.load(); }

}  // extern "C"
