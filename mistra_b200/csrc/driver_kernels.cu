// The layer loop of SUBROUTINE kpp_driver on the device (include/mistra_driver.h): CUDA kernels + C-ABI entries.
// Role in the reference: /root/reference/src/kpp.f90:4305-4470 - per layer, the scalars every *_drive call receives,
// the cloud bookkeeping, the advection source, and the choice of gas / aer / tot.
//
// Mapping: driver_layer_kernel gives 64 threads to a layer (four layers per block): thread i < nphrxn averages the
// photolysis rate i of the two bounding levels (coalesced rows of photol_j), thread 0 forms the scalars and the
// switches and writes the layer's mechanism.  The reference dispatches layer by layer; here the layers of each
// mechanism are gathered into ascending lists in three small launches (per-block counts by ballot, an exclusive scan
// over the blocks, an ordered write), so that the batch entries downstream see each mechanism as one batch.
// HBM-bound and tiny beside the integration: ~1 kB read and ~0.7 kB written per layer.
#include "../../include/mistra_driver.h"
#include "../../include/mistra_kpp.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

int mistra_internal_fail(int code, const std::string &msg);  // kpp_api.cu

namespace {

constexpr int DRV_LPB = 4;          // layers per block of driver_layer_kernel
constexpr int DRV_LIST = 1024;      // layers per block of the list kernels

__global__ void __launch_bounds__(256)
clip_negative_kernel(double *__restrict__ x, size_t count)                    // where (s1 < 0.d0) s1 = 0, kpp.f90:4305-4306
{
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256)
    if (x[i] < 0.0) x[i] = 0.0;
}

__global__ void __launch_bounds__(64 * DRV_LPB)
driver_layer_kernel(mistra_driver_args a, long long nlayer)
{
  const long long L = (long long)blockIdx.x * DRV_LPB + (threadIdx.x >> 6);
  const int i = threadIdx.x & 63;
  if (L >= nlayer) return;
  const int n = a.n, nkc = a.nkc;
  const long long col = L / n;
  const int k = (int)(L - col * n) + 1;                                       // the reference's level index
  const int n_min = a.box ? a.n_bl : 2, n_max = a.box ? a.n_bl : n - 1;
  if (k < n_min || k > n_max) {
    if (i == 0) a.mech[L] = -1;
    return;
  }
  // photolysis rates of the layer, kpp.f90:4344-4358
  const double u0min = a.lpBuys13_0D ? 1.75e-2 : 3.48e-2;
  if (i < a.nphrxn) {
    double ph = 0.0;
    if (a.u0[col] >= u0min) ph = (a.photol_j[(L - 1) * a.nphrxn + i] + a.photol_j[L * a.nphrxn + i]) / 2.0;
    a.ph_rat[L * a.nphrxn + i] = ph;
  }
  if (i != 0) return;
  const double airmolec = 6.022e+20 / 18.0;                                   // 4278
  const double Avogadro = 6.022140857e+23, conv1 = Avogadro * 1.e-6;          // constants.f90:36, 45
  const double te = a.t[L], air_cc = a.cm3[L], air = a.am3[L];                // 4315-4321
  const double h2o = a.xm1[L] * a.rho[L] / 1.8e-2;
  const double h2o_cc = a.xm1[L] * airmolec * a.rho[L];
  const double h2oppm = h2o_cc * 1.e6 / air_cc;
  const double pk = a.p[L];
  double cvv[4] = {0.0, 0.0, 0.0, 0.0}, xliq[4] = {0.0, 0.0, 0.0, 0.0};
  for (int kc = 0; kc < nkc && kc < 4; ++kc) {
    cvv[kc] = a.conv2[L * nkc + kc];                                          // 4327-4330
    xliq[kc] = 1.0;                                                           // 4374-4390
    if (k >= a.nf) xliq[kc] = 0.0;
    else if (a.cm[L * nkc + kc] == 0.0) xliq[kc] = 0.0;
    a.cloud[L * nkc + kc] = xliq[kc] == 1.0 ? 1 : 0;                          // 4392-4412: cloud follows xliq
  }
  double xhal = 1.0, xiod = 1.0;                                              // 4365-4371
  if (!a.halo) { xhal = 0.0; xiod = 0.0; }
  if (!a.iod) xiod = 0.0;
  double xhet1 = 1.0, xhet2 = 1.0;                                            // 4435-4438
  if (xliq[0] == 1.0) xhet1 = 0.0;
  if (xliq[1] == 1.0) xhet2 = 0.0;
  if (a.neula == 0 && k <= a.kinv && a.s1)                                    // 4441-4449
    for (int j = 0; j < a.nadv; ++j) {
      const int r = a.adv_row[j];
      if (r >= 0) a.s1[L * a.j1 + r] = a.s1[L * a.j1 + r] + a.xadv[j] * a.dt_ch * air / 86400.;
    }
  int mech = 0;                                                               // 4452-4468
  if (xliq[0] == 1.0 || xliq[1] == 1.0) mech = (xliq[2] == 1.0 || xliq[3] == 1.0) ? 2 : 1;
  else { xhet1 = 1.; xhet2 = 1.; }
  double *cb = a.cb1 + L * 4, *sc = a.scal + L * 13, *cv = a.cvv + L * 4;
  cb[0] = air_cc; cb[1] = te; cb[2] = h2oppm; cb[3] = pk;
  sc[0] = conv1; sc[1] = xhal; sc[2] = xiod; sc[3] = xhet1; sc[4] = xhet2;
  for (int kc = 0; kc < 4; ++kc) { sc[5 + kc] = xliq[kc]; sc[9 + kc] = cvv[kc]; cv[kc] = cvv[kc]; }
  a.air[L] = air;
  a.h2o[L] = h2o;
  a.mech[L] = mech;
}

// lists: per block of DRV_LIST layers the number of layers of each mechanism ...
__global__ void __launch_bounds__(DRV_LIST)
driver_count_kernel(const int32_t *__restrict__ mech, long long nlayer, int *__restrict__ blk_cnt)
{
  __shared__ int s_c[3];
  if (threadIdx.x < 3) s_c[threadIdx.x] = 0;
  __syncthreads();
  const long long L = (long long)blockIdx.x * DRV_LIST + threadIdx.x;
  const int m = L < nlayer ? mech[L] : -1;
  for (int x = 0; x < 3; ++x) {
    const unsigned b = __ballot_sync(0xffffffffu, m == x);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_c[x], __popc(b));
  }
  __syncthreads();
  if (threadIdx.x < 3) blk_cnt[blockIdx.x * 3 + threadIdx.x] = s_c[threadIdx.x];
}

// ... their exclusive scan over the blocks (one block; thread x < 3 walks mechanism x - the block count is small) ...
__global__ void driver_scan_kernel(int nblk, const int *__restrict__ blk_cnt, long long *__restrict__ blk_off,
                                   int64_t *__restrict__ count)
{
  const int x = threadIdx.x;
  if (x >= 3) return;
  long long run = 0;
  for (int b = 0; b < nblk; ++b) {
    blk_off[(size_t)b * 3 + x] = run;
    run += blk_cnt[(size_t)b * 3 + x];
  }
  count[x] = run;
}

// ... and the ordered write
__global__ void __launch_bounds__(DRV_LIST)
driver_write_kernel(const int32_t *__restrict__ mech, long long nlayer, const long long *__restrict__ blk_off,
                    int64_t *__restrict__ layers)
{
  __shared__ int s_w[3][DRV_LIST / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long L = (long long)blockIdx.x * DRV_LIST + threadIdx.x;
  const int m = L < nlayer ? mech[L] : -1;
  unsigned mine = 0;
  for (int x = 0; x < 3; ++x) {
    const unsigned b = __ballot_sync(0xffffffffu, m == x);
    if (lane == 0) s_w[x][warp] = __popc(b);
    if (m == x) mine = b;
  }
  __syncthreads();
  if (m < 0) return;
  long long off = blk_off[(size_t)blockIdx.x * 3 + m];
  for (int w = 0; w < warp; ++w) off += s_w[m][w];
  layers[(size_t)m * nlayer + off + __popc(mine & ((1u << lane) - 1u))] = L;
}

std::recursive_mutex g_mu;
std::atomic<long long> g_launches{0};
struct Scratch { char *p = nullptr; size_t bytes = 0; };
Scratch g_stage[16], g_work[16];

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

int check(int64_t ncol, const mistra_driver_args *a)
{
  if (ncol < 0 || !a) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad arguments");
  if (a->n < 3 || a->nf < 1 || a->nkc < 1 || a->nkc > 4 || a->nphrxn < 1 || a->nphrxn > 64)
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad sizes (n >= 3, 1 <= nkc <= 4, 1 <= nphrxn <= 64)");
  if (a->box && (a->n_bl < 2 || a->n_bl > a->n)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad n_bl");
  if (a->nadv < 0 || (a->nadv > 0 && a->neula == 0 && a->s1 && (!a->adv_row || !a->xadv || a->j1 < 1)))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad advection list");
  if ((a->s1 && a->j1 < 1) || (a->s3 && a->j5 < 1)) return mistra_internal_fail(MISTRA_KPP_EINVAL, "bad j1 / j5");
  if (!a->count) return mistra_internal_fail(MISTRA_KPP_EINVAL, "null count");
  if (ncol > 0 && (!a->u0 || !a->t || !a->p || !a->rho || !a->cm3 || !a->am3 || !a->xm1 || !a->conv2 || !a->cm || !a->cloud ||
                   !a->photol_j || !a->cb1 || !a->scal || !a->ph_rat || !a->air || !a->h2o || !a->cvv || !a->mech || !a->layers))
    return mistra_internal_fail(MISTRA_KPP_EINVAL, "null array");
  return 0;
}

}  // namespace

extern "C" {

int mistra_driver_layers_device(int64_t ncol, const mistra_driver_args *d_a, void *stream)
{
  if (int rc = check(ncol, d_a)) return rc;
  const mistra_driver_args &a = *d_a;
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  if (ncol == 0) { CKW(cudaMemsetAsync(a.count, 0, 3 * sizeof(int64_t), st)); return 0; }
  const long long nlayer = (long long)ncol * a.n;
  if (nlayer > (1ll << 31) - DRV_LIST) return mistra_internal_fail(MISTRA_KPP_EINVAL, "too many layers for one call");
  int nl = 0;
  if (a.s1) { clip_negative_kernel<<<2048, 256, 0, st>>>(a.s1, (size_t)nlayer * a.j1); CKW(cudaGetLastError()); ++nl; }
  if (a.s3) { clip_negative_kernel<<<2048, 256, 0, st>>>(a.s3, (size_t)nlayer * a.j5); CKW(cudaGetLastError()); ++nl; }
  driver_layer_kernel<<<(unsigned)((nlayer + DRV_LPB - 1) / DRV_LPB), 64 * DRV_LPB, 0, st>>>(a, nlayer);
  CKW(cudaGetLastError());
  const int nblk = (int)((nlayer + DRV_LIST - 1) / DRV_LIST);
  const size_t o_off = ((size_t)nblk * 3 * sizeof(int) + 255) & ~(size_t)255;
  if (int rc = grow(g_work[dev], o_off + (size_t)nblk * 3 * sizeof(long long))) return rc;
  int *blk_cnt = (int *)g_work[dev].p;
  long long *blk_off = (long long *)(g_work[dev].p + o_off);
  driver_count_kernel<<<nblk, DRV_LIST, 0, st>>>(a.mech, nlayer, blk_cnt);
  CKW(cudaGetLastError());
  driver_scan_kernel<<<1, 32, 0, st>>>(nblk, blk_cnt, blk_off, a.count);
  CKW(cudaGetLastError());
  driver_write_kernel<<<nblk, DRV_LIST, 0, st>>>(a.mech, nlayer, blk_off, a.layers);
  CKW(cudaGetLastError());
  g_launches.fetch_add(nl + 4);
  return 0;
}

int mistra_driver_layers(int64_t ncol, const mistra_driver_args *a, void *stream)
{
  if (int rc = check(ncol, a)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  int dev = -1;
  CKW(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return mistra_internal_fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  const size_t nc = (size_t)ncol, n = a->n, nl = nc * n, nkc = a->nkc, np = a->nphrxn;
  struct Item { const void *h; size_t bytes; bool in, out; void **slot; };
  mistra_driver_args d = *a;
  std::vector<Item> items = {
      {a->u0, nc * 8, true, false, (void **)&d.u0}, {a->t, nl * 8, true, false, (void **)&d.t},
      {a->p, nl * 8, true, false, (void **)&d.p}, {a->rho, nl * 8, true, false, (void **)&d.rho},
      {a->cm3, nl * 8, true, false, (void **)&d.cm3}, {a->am3, nl * 8, true, false, (void **)&d.am3},
      {a->xm1, nl * 8, true, false, (void **)&d.xm1}, {a->conv2, nl * nkc * 8, true, false, (void **)&d.conv2},
      {a->cm, nl * nkc * 8, true, false, (void **)&d.cm}, {a->cloud, nl * nkc * 4, true, true, (void **)&d.cloud},
      {a->photol_j, nl * np * 8, true, false, (void **)&d.photol_j},
      {a->adv_row, (size_t)a->nadv * 4, true, false, (void **)&d.adv_row}, {a->xadv, (size_t)a->nadv * 8, true, false, (void **)&d.xadv},
      {a->s1, a->s1 ? nl * (size_t)a->j1 * 8 : 0, true, true, (void **)&d.s1},
      {a->s3, a->s3 ? nl * (size_t)a->j5 * 8 : 0, true, true, (void **)&d.s3},
      {a->cb1, nl * 4 * 8, true, true, (void **)&d.cb1}, {a->scal, nl * 13 * 8, true, true, (void **)&d.scal},
      {a->ph_rat, nl * np * 8, true, true, (void **)&d.ph_rat}, {a->air, nl * 8, true, true, (void **)&d.air},
      {a->h2o, nl * 8, true, true, (void **)&d.h2o}, {a->cvv, nl * 4 * 8, true, true, (void **)&d.cvv},
      {a->mech, nl * 4, false, true, (void **)&d.mech}, {a->layers, 3 * nl * 8, true, true, (void **)&d.layers},
      {a->count, 3 * 8, false, true, (void **)&d.count}};
  size_t total = 0;
  for (auto &it : items) total += (it.bytes + 255) & ~(size_t)255;
  if (int rc = grow(g_stage[dev], total + 256)) return rc;
  char *p = g_stage[dev].p;
  for (auto &it : items) {
    if (!it.h || it.bytes == 0) { *it.slot = nullptr; continue; }
    *it.slot = p;
    if (it.in) CKW(cudaMemcpyAsync(p, it.h, it.bytes, cudaMemcpyHostToDevice, st));
    p += (it.bytes + 255) & ~(size_t)255;
  }
  if (!d.count) d.count = (int64_t *)p;
  int rc = mistra_driver_layers_device(ncol, &d, stream);
  if (rc) { cudaStreamSynchronize(st); return rc; }
  for (auto &it : items)
    if (it.out && it.h && it.bytes) CKW(cudaMemcpyAsync((void *)it.h, *it.slot, it.bytes, cudaMemcpyDeviceToHost, st));
  CKW(cudaStreamSynchronize(st));
  return 0;
}

int64_t mistra_driver_launch_count(void) { return g_launches.load(); }

}  // extern "C"
