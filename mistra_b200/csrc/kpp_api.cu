// C-ABI layer of libmistra_kpp.so (include/mistra_kpp.h): option decoding as in
// Rosenbrock_x (/root/reference/src/gas.f:950-1051), device workspaces, host
// staging and kernel launches.  No CPU fallback: every compute entry needs a
// CUDA device.
#include "../../include/mistra_kpp.h"
#include "kpp_onchip.h"
#include "../../include/mistra_kpp_rates.h"
#include "../../include/mistra_rconst_cuda.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

std::mutex g_mu;                 // the device table, the kernel-variant choice
thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<long long> g_launches_variant[2];
void *g_oc_aux = nullptr;      // development aid of the on-chip kernels (OC_DEBUG builds dump checkpoints here)
long long g_oc_flags = 0;

struct MechState {
  // variant 0: one cell per thread (ros3_kernel.inc)
  double *ws = nullptr;    // lane workspace of kernels on the caller's stream
  double *ws2 = nullptr;   // second workspace: odd chunks of the host-buffer pipeline (allocated on first use)
  size_t ws_bytes = 0;
  int blocks = 0;
  int coef_variant = -1;   // -1 unset, 0 f64 literals, 1 f32 literals
  // variant 1: on-chip kernel (ros3_onchip.inc)
  unsigned short *oc_tab = nullptr;   // device copy of the instruction streams
  double *oc_ws = nullptr, *oc_ws2 = nullptr;   // strict build only: tail block of the LU factors per cell slot
  size_t oc_ws_bytes = 0;
  int oc_blocks = 0;
  int oc_lit_variant = -1;
};

struct DeviceState {
  bool init = false;
  int dev = -1;
  int num_sm = 0;
  cudaStream_t stream = nullptr;
  unsigned long long *counter = nullptr;   // [3]: one cell counter per pipeline slot, [2]: the hand-off pass
  char *handoff = nullptr;                 // deferred-cell count and list, (T, H) per cell, spare statistics of a call
  size_t handoff_bytes = 0;
  char *fix_dev = nullptr, *fix_host = nullptr;   // host-buffer entries: results of the handed-over cells, compact (device / pinned)
  size_t fix_bytes = 0;
  MechState mech[3];
  // device staging for the host-buffer entry
  void *d_stage = nullptr;
  size_t d_stage_bytes = 0;
  void *d_rates = nullptr;            // scratch of mistra_kpp_integrate_rates (expanded rate arrays, two slots)
  size_t d_rates_bytes = 0;
  // copy streams / events of the chunk pipeline of the host-buffer entry
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_k2 = nullptr;
  cudaEvent_t ev_h[64] = {}, ev_k[64] = {}, ev_start = nullptr, ev_join = nullptr;
  // last kernel of each slot: a launch on another stream waits for it (the slot's workspace and
  // cell counter must not be shared by two running kernels)
  cudaEvent_t ev_slot[2] = {};
};

#ifndef KPP_ONCHIP_MAX_CELLS_G
#define KPP_ONCHIP_MAX_CELLS_G 8192
#endif
#ifndef KPP_ONCHIP_MAX_CELLS_A
#define KPP_ONCHIP_MAX_CELLS_A 98304
#endif
constexpr int kMaxDev = 16;
DeviceState g_dev[kMaxDev];
// one lock per device: host threads that drive different GPUs of one process run concurrently
std::mutex g_dev_mu[kMaxDev];

int fail(int code, const std::string &msg)
{
  g_err = msg;
  return code;
}
}  // namespace

// shared with the other translation units of the library (bins_kernels.cu)
int mistra_internal_fail(int code, const std::string &msg) { return fail(code, msg); }

namespace {

int cuda_fail(cudaError_t e, const char *what)
{
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? MISTRA_KPP_ENODEVICE
         : (e == cudaErrorMemoryAllocation)                           ? MISTRA_KPP_ENOMEM
                                                                      : MISTRA_KPP_ECUDA;
}

#define CK(call)                                          \
  do {                                                    \
    cudaError_t e_ = (call);                              \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);   \
  } while (0)

const KppMechInfo *mech_info(int mech)
{
  switch (mech) {
    case MISTRA_KPP_GAS: return kpp_mech_info_g();
    case MISTRA_KPP_AER: return kpp_mech_info_a();
    case MISTRA_KPP_TOT: return kpp_mech_info_t();
  }
  return nullptr;
}

int get_device(DeviceState **out)
{
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxDev) return fail(MISTRA_KPP_ENODEVICE, "device index out of range");
  DeviceState &d = g_dev[dev];
  if (!d.init) {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    d.dev = dev;
    d.num_sm = p.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&d.counter, 3 * sizeof(unsigned long long)));
    for (auto &e : d.ev_slot) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    d.init = true;
  }
  *out = &d;
  return 0;
}

// Rosenbrock_x option decoding, gas.f:950-1051 (same tests, same defaults)
int decode_opts(const mistra_kpp_opts *o, double t0, double t1, KppBatch *b)
{
  mistra_kpp_opts dflt;
  if (!o) { mistra_kpp_default_opts(&dflt); o = &dflt; }
  const double Roundoff = 2.220446049250313e-16, DeltaMin = 1.0e-5;
  if (o->max_steps == 0) b->max_steps = 100000;
  else if (o->max_steps > 0) b->max_steps = o->max_steps;
  else return fail(MISTRA_KPP_EOPTS, "max_steps < 0 (IPAR(3), ros_ErrorMsg -1)");
  if (o->hmin == 0.0) b->hmin = 0.0;
  else if (o->hmin > 0.0) b->hmin = o->hmin;
  else return fail(MISTRA_KPP_EOPTS, "hmin < 0 (RPAR(1), ros_ErrorMsg -3)");
  if (o->hmax == 0.0) b->hmax = std::fabs(t1 - t0);
  else if (o->hmax > 0.0) b->hmax = std::fmin(std::fabs(o->hmax), std::fabs(t1 - t0));
  else return fail(MISTRA_KPP_EOPTS, "hmax < 0 (RPAR(2), ros_ErrorMsg -3)");
  if (o->hstart == 0.0) b->hstart = std::fmax(b->hmin, DeltaMin);
  else if (o->hstart > 0.0) b->hstart = std::fmin(std::fabs(o->hstart), std::fabs(t1 - t0));
  else return fail(MISTRA_KPP_EOPTS, "hstart < 0 (RPAR(3), ros_ErrorMsg -3)");
  auto fac = [&](double v, double d, double *out, const char *nm) -> int {
    if (v == 0.0) *out = d;
    else if (v > 0.0) *out = v;
    else return fail(MISTRA_KPP_EOPTS, std::string(nm) + " < 0 (RPAR(4..7), ros_ErrorMsg -4)");
    return 0;
  };
  int rc;
  if ((rc = fac(o->facmin, 0.2, &b->facmin, "facmin"))) return rc;
  if ((rc = fac(o->facmax, 6.0, &b->facmax, "facmax"))) return rc;
  if ((rc = fac(o->facrej, 0.1, &b->facrej, "facrej"))) return rc;
  if ((rc = fac(o->facsafe, 0.9, &b->facsafe, "facsafe"))) return rc;
  if (!(o->atol > 0.0) || !(o->rtol > 10.0 * Roundoff) || !(o->rtol < 1.0))
    return fail(MISTRA_KPP_EOPTS, "unreasonable tolerances (ros_ErrorMsg -5)");
  b->atol = o->atol;
  b->rtol = o->rtol;
  b->autonomous = o->autonomous ? 1 : 0;
  b->t0 = t0;
  b->t1 = t1;
  return 0;
}

double literal_value(const char *lit, int f32)
{
  // Fortran default-REAL literal: binary32 under the reference's preferred flags
  return f32 ? (double)strtof(lit, nullptr) : strtod(lit, nullptr);
}

// Kernel variant of a mechanism: 0 = one cell per thread with the lane workspace in HBM (ros3_kernel.inc), 1 = the
// on-chip kernel (ros3_onchip.inc: one persistent block per SM, LU in shared memory / registers, DRAM traffic =
// the compulsory I/O), -1 = by batch size (the default): the on-chip kernel keeps only a few hundred cells in
// flight, so it is the faster one for small batches and for the one-cell-per-call box model, the cell-per-thread
// kernel needs tens of thousands of cells to fill the device and is the faster one for large batches (measured
// crossovers below, tools/variant_sweep.py).  mistra_kpp_set_kernel() pins a variant; MISTRA_KPP_ONCHIP=0/1 in the
// environment does the same for every mechanism.  Both are CUDA paths, there is no CPU path.
int g_variant[3] = {-1, -1, -1};
const int64_t kOnchipMaxCells[3] = {KPP_ONCHIP_MAX_CELLS_G, KPP_ONCHIP_MAX_CELLS_A, 0};
bool want_onchip(const KppMechInfo *mi, int mech, int64_t ncell)
{
  if (!mi->oc) return false;
  int v = g_variant[mech];
  if (v < 0) {
    const char *e = getenv("MISTRA_KPP_ONCHIP");
    if (e && (atoi(e) == 0 || atoi(e) == 1)) v = atoi(e);
  }
  if (v >= 0) return v == 1;
  return ncell <= kOnchipMaxCells[mech];
}
// Hand-off of long cells (kpp_batch.h): step attempts a cell may make in the cell-per-thread kernel before the on-chip
// kernel continues it.  -1 = default: 12 for aer when the library chose the kernel by batch size (a steady-state cell
// needs 7 - 8 attempts; a cell with rejected steps or in a transient needs tens to hundreds and would hold a lane at
// 6.5 ms per step, the on-chip kernel runs it at 0.15 ms per step), 0 otherwise and for gas / tot.
int g_handoff[3] = {-1, -1, -1};
int handoff_steps(const KppMechInfo *mi, int mech)
{
  if (!mi->oc) return 0;
  if (g_handoff[mech] >= 0) return g_handoff[mech];
  if (const char *e = getenv("MISTRA_KPP_HANDOFF")) return mech == 1 && atoi(e) > 0 ? atoi(e) : 0;
  if (g_variant[mech] >= 0 || getenv("MISTRA_KPP_ONCHIP")) return 0;
  return mech == 1 ? 12 : 0;
}
int ensure_mech(DeviceState &d, int mech, const KppMechInfo *mi, int f32, cudaStream_t st, bool onchip)
{
  MechState &ms = d.mech[mech];
  if (onchip) {
    if (!ms.oc_tab) {
      int per_sm = 0;
      CK(cudaFuncSetAttribute(mi->oc->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mi->oc->smem_bytes));
      CK(cudaFuncSetAttribute(mi->oc->kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mi->oc->kernel, mi->oc->threads, mi->oc->smem_bytes));
      if (per_sm < 1) return fail(MISTRA_KPP_ECUDA, "on-chip kernel does not fit on an SM");
      ms.oc_blocks = d.num_sm;      // one persistent block per SM
      CK(cudaMalloc(&ms.oc_tab, mi->oc->table_count * sizeof(unsigned short)));
      CK(cudaMemcpyAsync(ms.oc_tab, mi->oc->tables, mi->oc->table_count * sizeof(unsigned short),
                         cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
#ifdef KPP_STRICT
      // strict build: the tail block of the LU factors is also written out, for the reference-order
      // backward substitution
      ms.oc_ws_bytes = (size_t)ms.oc_blocks * mi->oc->slots * mi->oc->tail * mi->oc->tail * sizeof(double);
      CK(cudaMalloc(&ms.oc_ws, ms.oc_ws_bytes));
#endif
    }
    if (ms.oc_lit_variant != f32) {
      double h[64];
      if (mi->oc->nlit > 64) return fail(MISTRA_KPP_EINVAL, "coefficient table too large");
      for (int i = 0; i < mi->oc->nlit; ++i) h[i] = literal_value(mi->oc->literals[i], f32);
      CK(cudaDeviceSynchronize());    // kernels that still read the previous table
      CK(mi->oc->set_lit(h, st));
      CK(cudaStreamSynchronize(st));  // h is on the stack
      ms.oc_lit_variant = f32;
    }
    return 0;
  }
  if (!ms.ws) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mi->kernel, KPP_BLOCK, 0));
    if (per_sm < 1) return fail(MISTRA_KPP_ECUDA, "kernel does not fit on an SM");
    if (const char *e = getenv("MISTRA_KPP_BLOCKS_PER_SM")) {
      int v = atoi(e);
      if (v >= 1 && v < per_sm) per_sm = v;
    }
    ms.blocks = per_sm * d.num_sm;
    const size_t warps = (size_t)ms.blocks * (KPP_BLOCK / 32);
    ms.ws_bytes = warps * (size_t)mi->nslot * 32 * sizeof(double);
    CK(cudaMalloc(&ms.ws, ms.ws_bytes));
  }
  if (ms.coef_variant != f32) {
    double h[64];
    if (mi->ncoef > 64) return fail(MISTRA_KPP_EINVAL, "coefficient table too large");
    for (int i = 0; i < mi->ncoef; ++i) h[i] = literal_value(mi->coef_literals[i], f32);
    CK(mi->set_coef(h, st));
    CK(cudaStreamSynchronize(st));  // h is on the stack
    ms.coef_variant = f32;
  }
  return 0;
}

// Hand-off context of one mistra_kpp_integrate_device call (kpp_batch.h): arrays over the cells of the whole call.
struct Handoff {
  int soft;
  unsigned long long *count;
  long long *list;
  double *cont;
  int32_t *stats;      // used when the caller keeps no statistics
};

int handoff_begin(DeviceState &d, int64_t ncell, int soft, Handoff *ho, cudaStream_t st)
{
  // [count | list: ncell x 8 | cont: ncell x 16 | statistics: ncell x 32]
  const size_t o_list = 256, o_cont = o_list + (((size_t)ncell * 8 + 255) & ~(size_t)255),
               o_stats = o_cont + (((size_t)ncell * 16 + 255) & ~(size_t)255), need = o_stats + (size_t)ncell * 32;
  if (d.handoff_bytes < need) {
    if (d.handoff) { CK(cudaDeviceSynchronize()); cudaFree(d.handoff); d.handoff = nullptr; d.handoff_bytes = 0; }
    CK(cudaMalloc(&d.handoff, need));
    d.handoff_bytes = need;
  }
  ho->soft = soft;
  ho->count = (unsigned long long *)d.handoff;
  ho->list = (long long *)(d.handoff + o_list);
  ho->cont = (double *)(d.handoff + o_cont);
  ho->stats = (int32_t *)(d.handoff + o_stats);
  CK(cudaStreamWaitEvent(st, d.ev_slot[0], 0));         // the hand-off pass of a call on another stream may still read them
  CK(cudaMemsetAsync(ho->count, 0, sizeof(unsigned long long), st));
  return 0;
}

int launch_device(DeviceState &d, int mech, int64_t ncell, const double *d_rconst,
                  const double *d_fix, double *d_var, double t0, double t1,
                  const mistra_kpp_opts *o, int32_t *d_ierr, int32_t *d_stats, double *d_hexit,
                  double *d_texit, cudaStream_t st, int slot, bool onchip, const Handoff *ho = nullptr, int64_t off = 0)
{
  const KppMechInfo *mi = mech_info(mech);
  KppBatch b;
  memset(&b, 0, sizeof(b));
  int rc = decode_opts(o, t0, t1, &b);
  if (rc) return rc;
  const int f32 = o ? (o->f32_literals ? 1 : 0) : 1;
  if ((rc = ensure_mech(d, mech, mi, f32, st, onchip))) return rc;
  if (ncell == 0) return 0;
  MechState &ms = d.mech[mech];
  b.rconst = d_rconst;
  b.fix = d_fix;
  b.var = d_var;
  b.ierr = d_ierr;
  b.stats = d_stats;
  b.hexit = d_hexit;
  b.texit = d_texit;
  b.ncell = ncell;
  if (onchip) {
    if (slot == 1 && !ms.oc_ws2 && ms.oc_ws_bytes) CK(cudaMalloc(&ms.oc_ws2, ms.oc_ws_bytes));
    b.ws = slot ? ms.oc_ws2 : ms.oc_ws;
  } else {
    if (slot == 1 && !ms.ws2 && ms.ws_bytes) CK(cudaMalloc(&ms.ws2, ms.ws_bytes));
    b.ws = slot ? ms.ws2 : ms.ws;
  }
  b.oc_tab = ms.oc_tab;
  b.oc_aux = g_oc_aux;
  b.oc_flags = g_oc_flags;
  b.counter = d.counter + slot;
  CK(cudaStreamWaitEvent(st, d.ev_slot[slot], 0));
  CK(cudaMemsetAsync(b.counter, 0, sizeof(unsigned long long), st));
  if (ho && !onchip) {                                 // long cells are retired at a step boundary (kpp_batch.h)
    b.soft_steps = ho->soft;
    b.defer_count = ho->count;
    b.defer_list = ho->list;
    b.cell_base = off;
    b.cont = ho->cont + 2 * off;
    if (!b.stats) b.stats = ho->stats + 8 * off;
  }
  if (onchip) {
    const long long need = (ncell + mi->oc->slots - 1) / mi->oc->slots;
    int blocks = (int)(need < ms.oc_blocks ? need : ms.oc_blocks);
    CK(mi->oc->launch(b, blocks, st));
  } else {
    long long need_blocks = (ncell + KPP_BLOCK - 1) / KPP_BLOCK;
    int blocks = (int)(need_blocks < ms.blocks ? need_blocks : ms.blocks);
    CK(mi->launch(b, blocks, st));
  }
  CK(cudaEventRecord(d.ev_slot[slot], st));
  g_launches.fetch_add(1);
  g_launches_variant[onchip ? 1 : 0].fetch_add(1);
  return 0;
}


// The cells the cell-per-thread launches of a call have retired unfinished, continued by the on-chip kernel: one
// launch at the end of the call (the on-chip blocks need whole SMs, which the chunks of pass 1 only free at their end).
int launch_handoff_pass(DeviceState &d, int mech, int64_t ncell, const double *d_rconst, const double *d_fix, double *d_var,
                        double t0, double t1, const mistra_kpp_opts *o, int32_t *d_ierr, int32_t *d_stats, double *d_hexit,
                        double *d_texit, cudaStream_t st, const Handoff &ho)
{
  const KppMechInfo *mi = mech_info(mech);
  KppBatch b;
  memset(&b, 0, sizeof(b));
  int rc = decode_opts(o, t0, t1, &b);
  if (rc) return rc;
  MechState &ms = d.mech[mech];
  b.rconst = d_rconst; b.fix = d_fix; b.var = d_var;
  b.ierr = d_ierr; b.stats = d_stats ? d_stats : ho.stats; b.hexit = d_hexit; b.texit = d_texit;
  b.ncell = ncell;
  b.ws = ms.oc_ws;
  b.oc_tab = ms.oc_tab; b.oc_aux = g_oc_aux; b.oc_flags = g_oc_flags;
  b.resume = 1;
  b.cont = ho.cont;
  b.list = ho.list;
  b.list_count = ho.count;
  b.counter = d.counter + 2;
  CK(cudaMemsetAsync(b.counter, 0, sizeof(unsigned long long), st));
  const long long need = (ncell + mi->oc->slots - 1) / mi->oc->slots;
  CK(mi->oc->launch(b, (int)(need < ms.oc_blocks ? need : ms.oc_blocks), st));
  CK(cudaEventRecord(d.ev_slot[0], st));
  g_launches.fetch_add(1);
  g_launches_variant[1].fetch_add(1);
  return 0;
}

// Host-buffer entries: the per-chunk copies have taken the handed-over cells home in their intermediate state; their
// final rows are gathered into a compact block, copied once and put in place on the host.
__global__ void handoff_gather_kernel(const long long *__restrict__ list, long long count, int nvar, const double *__restrict__ var,
                                      const int32_t *__restrict__ ierr, const int32_t *__restrict__ stats,
                                      const double *__restrict__ hexit, const double *__restrict__ texit,
                                      double *__restrict__ cvar, double *__restrict__ chx, double *__restrict__ ctx,
                                      int32_t *__restrict__ cst, int32_t *__restrict__ cie)
{
  for (long long i = blockIdx.x; i < count; i += gridDim.x) {
    const long long c = list[i];
    for (int j = threadIdx.x; j < nvar; j += blockDim.x) cvar[i * nvar + j] = var[c * nvar + j];
    if (threadIdx.x < 8 && stats) cst[i * 8 + threadIdx.x] = stats[c * 8 + threadIdx.x];
    if (threadIdx.x == 0) {
      if (ierr) cie[i] = ierr[c];
      if (hexit) chx[i] = hexit[c];
      if (texit) ctx[i] = texit[c];
    }
  }
}

// st has been synchronised (pass 2 done) and so has the D2H stream (the per-chunk rows are home).
int handoff_fixup_host(DeviceState &d, const KppMechInfo *mi, int64_t ncell, const Handoff &ho, const double *d_var,
                       const int32_t *d_ie, const int32_t *d_st, const double *d_hx, const double *d_tx, double *var,
                       int32_t *ierr, int32_t *stats, double *hexit, double *texit, cudaStream_t st)
{
  unsigned long long cnt = 0;
  CK(cudaMemcpyAsync(&cnt, ho.count, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (cnt == 0) return 0;
  const size_t n = (size_t)cnt, nvar = (size_t)mi->nvar;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_hx = al(n * nvar * 8), o_tx = o_hx + al(n * 8), o_st = o_tx + al(n * 8), o_ie = o_st + al(n * 32),
               o_ls = o_ie + al(n * 4), total = o_ls + al(n * 8);
  if (d.fix_bytes < total) {
    if (d.fix_dev) cudaFree(d.fix_dev);
    if (d.fix_host) cudaFreeHost(d.fix_host);
    d.fix_dev = d.fix_host = nullptr; d.fix_bytes = 0;
    CK(cudaMalloc(&d.fix_dev, total));
    CK(cudaHostAlloc(&d.fix_host, total, cudaHostAllocPortable));
    d.fix_bytes = total;
  }
  char *g = d.fix_dev;
  handoff_gather_kernel<<<(unsigned)(n < 4096 ? n : 4096), 128, 0, st>>>(
      ho.list, (long long)cnt, mi->nvar, d_var, ierr ? d_ie : nullptr, stats ? d_st : nullptr, hexit ? d_hx : nullptr,
      texit ? d_tx : nullptr, (double *)g, (double *)(g + o_hx), (double *)(g + o_tx), (int32_t *)(g + o_st), (int32_t *)(g + o_ie));
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(g + o_ls, ho.list, n * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(d.fix_host, g, total, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const char *h = d.fix_host;
  const long long *ls = (const long long *)(h + o_ls);
  const double *hv = (const double *)h, *hhx = (const double *)(h + o_hx), *htx = (const double *)(h + o_tx);
  const int32_t *hst = (const int32_t *)(h + o_st), *hie = (const int32_t *)(h + o_ie);
  for (size_t i = 0; i < n; ++i) {
    const long long c = ls[i];
    if (c < 0 || c >= ncell) return fail(MISTRA_KPP_ECUDA, "hand-off list holds a cell index outside the batch");
    memcpy(var + (size_t)c * nvar, hv + i * nvar, nvar * 8);
    if (ierr) ierr[c] = hie[i];
    if (stats) memcpy(stats + (size_t)c * 8, hst + i * 8, 32);
    if (hexit) hexit[c] = hhx[i];
    if (texit) texit[c] = htx[i];
  }
  return 0;
}

// ---- compact rate inputs -> the NSPEC-indexed arrays Update_RCONST_x reads (include/mistra_kpp_rates.h) ------------
__global__ void rates_expand_kernel(double *__restrict__ full, const double *__restrict__ val, const int *__restrict__ idx,
                                    long long nrow, int n, int nspec)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (row = cell * nk + k, j)
  if (i >= nrow * n) return;
  const long long row = i / n;
  const int j = (int)(i - row * n);
  full[row * nspec + idx[j]] = val[i];
}

__global__ void rates_conc_kernel(double *__restrict__ conc, const double *__restrict__ var, const double *__restrict__ fix,
                                  long long ncell, int nvar, int nfix)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nspec = nvar + nfix;
  if (i >= ncell * nspec) return;
  const long long c = i / nspec;
  const int s = (int)(i - c * nspec);
  conc[i] = s < nvar ? var[c * nvar + s] : fix[c * nfix + (s - nvar)];
}

int ensure_streams(DeviceState *d)
{
  if (d->s_h2d) return 0;
  CK(cudaStreamCreateWithFlags(&d->s_h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&d->s_d2h, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&d->s_k2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
  for (auto &e : d->ev_h) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : d->ev_k) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming));
  return 0;
}

// FP64 FMA throughput probe: 8 independent DFMA chains per thread, every lane busy.
// This is the measured denominator of the FP64 roofline (MEASURED_PEAKS.json has
// no FP64 entry); it is not part of the chemistry path.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace

extern "C" {

// Measured FP64 FMA peak of the current device in TFLOP/s (FMA = 2 flops); < 0 on error.
double mistra_kpp_fp64_peak_tflops(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceState *d;
  if (get_device(&d)) return -1.0;
  double *out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
  const int blocks = d->num_sm * 8, iters = 1 << 15;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = -1.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, d->stream);
    fp64_peak_kernel<<<blocks, 256, 0, d->stream>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, d->stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * 256.0 * blocks / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

void mistra_kpp_default_opts(mistra_kpp_opts *o)
{
  memset(o, 0, sizeof(*o));
  o->rtol = 1.0e-3;    // gas.f:745
  o->atol = 1.0e-25;   // gas.f:746
  o->hstart = 1.0e-3;  // gas.f:743
  o->f32_literals = 1;
}

int mistra_kpp_query(int mech, int *nvar, int *nfix, int *nreact, int *lu_nonzero)
{
  const KppMechInfo *mi = mech_info(mech);
  if (!mi) return fail(MISTRA_KPP_EINVAL, "unknown mechanism id");
  if (nvar) *nvar = mi->nvar;
  if (nfix) *nfix = mi->nfix;
  if (nreact) *nreact = mi->nreact;
  if (lu_nonzero) *lu_nonzero = mi->lu_nonzero;
  return 0;
}

const char *mistra_kpp_spc_name_impl(int mech, int i);  // kpp_names.cpp (generated table)

const char *mistra_kpp_spc_name(int mech, int i) { return mistra_kpp_spc_name_impl(mech, i); }

int mistra_kpp_integrate_device(int mech, int64_t ncell, const double *d_rconst,
                                const double *d_fix, double *d_var, double t0, double t1,
                                const mistra_kpp_opts *o, int32_t *d_ierr, int32_t *d_stats,
                                double *d_hexit, double *d_texit, void *stream)
{
  if (!mech_info(mech)) return fail(MISTRA_KPP_EINVAL, "unknown mechanism id");
  if (ncell < 0) return fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell > 0 && (!d_rconst || !d_fix || !d_var))
    return fail(MISTRA_KPP_EINVAL, "null rconst/fix/var");
  DeviceState *d;
  int rc;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    rc = get_device(&d);
  }
  if (rc) return rc;
  std::lock_guard<std::mutex> dl(g_dev_mu[d->dev]);
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the legacy default stream, as in the CUDA runtime
  // A long batch runs as a chain of chunks that alternate between two streams / workspaces (as the host-buffer entry
  // does): while the lanes of one chunk run out of cells the next chunk already has its blocks on the device, so the
  // SMs do not idle through the tail of one big launch while its slowest cells finish (aer, 588 000 cells: two halves
  // +3 %, sixteen chunks more; measured).  MISTRA_KPP_SPLIT=0 switches it off.
  static const bool split = !(getenv("MISTRA_KPP_SPLIT") && atoi(getenv("MISTRA_KPP_SPLIT")) == 0);
  const bool oc = want_onchip(mech_info(mech), mech, ncell);
  Handoff ho, *hop = nullptr;
  if (!oc && ncell > 0 && handoff_steps(mech_info(mech), mech) > 0) {
    if ((rc = ensure_mech(*d, mech, mech_info(mech), o ? (o->f32_literals ? 1 : 0) : 1, st, true))) return rc;
    if ((rc = handoff_begin(*d, ncell, handoff_steps(mech_info(mech), mech), &ho, st))) return rc;
    hop = &ho;
  }
  if (split && !oc && ncell >= 8LL * d->num_sm * 2 * KPP_BLOCK) {
    const KppMechInfo *mi = mech_info(mech);
    if ((rc = ensure_streams(d))) return rc;
    if ((rc = ensure_mech(*d, mech, mi, o ? (o->f32_literals ? 1 : 0) : 1, st, false))) return rc;
    const int64_t resident = (int64_t)d->mech[mech].blocks * KPP_BLOCK;
    int64_t nchunk = ncell / (2 * resident);
    if (nchunk < 2) nchunk = 2;
    if (nchunk > 16) nchunk = 16;
    const int64_t per = (ncell + nchunk - 1) / nchunk;
    CK(cudaEventRecord(d->ev_start, st));
    CK(cudaStreamWaitEvent(d->s_k2, d->ev_start, 0));
    for (int64_t c = 0, off = 0; off < ncell; ++c, off += per) {
      const int slot = (int)(c & 1);
      const int64_t m = (ncell - off) < per ? (ncell - off) : per;
      if ((rc = launch_device(*d, mech, m, d_rconst + off * mi->nreact, d_fix + off * mi->nfix, d_var + off * mi->nvar,
                              t0, t1, o, d_ierr ? d_ierr + off : nullptr, d_stats ? d_stats + 8 * off : nullptr,
                              d_hexit ? d_hexit + off : nullptr, d_texit ? d_texit + off : nullptr,
                              slot ? d->s_k2 : st, slot, false, hop, off)))
        return rc;
    }
    CK(cudaEventRecord(d->ev_join, d->s_k2));
    CK(cudaStreamWaitEvent(st, d->ev_join, 0));
    if (hop) return launch_handoff_pass(*d, mech, ncell, d_rconst, d_fix, d_var, t0, t1, o, d_ierr, d_stats, d_hexit, d_texit, st, ho);
    return 0;
  }
  if ((rc = launch_device(*d, mech, ncell, d_rconst, d_fix, d_var, t0, t1, o, d_ierr, d_stats,
                          d_hexit, d_texit, st, 0, oc, hop, 0)))
    return rc;
  if (hop) return launch_handoff_pass(*d, mech, ncell, d_rconst, d_fix, d_var, t0, t1, o, d_ierr, d_stats, d_hexit, d_texit, st, ho);
  return 0;
}

int mistra_kpp_integrate(int mech, int64_t ncell, const double *rconst, const double *fix,
                         double *var, double t0, double t1, const mistra_kpp_opts *o,
                         int32_t *ierr, int32_t *stats, double *hexit, double *texit,
                         void *stream)
{
  const KppMechInfo *mi = mech_info(mech);
  if (!mi) return fail(MISTRA_KPP_EINVAL, "unknown mechanism id");
  if (ncell < 0) return fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell > 0 && (!rconst || !fix || !var)) return fail(MISTRA_KPP_EINVAL, "null rconst/fix/var");
  {
    KppBatch tmp;
    int rc = decode_opts(o, t0, t1, &tmp);  // reject bad options before touching the device
    if (rc) return rc;
  }
  DeviceState *d;
  int rc;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    rc = get_device(&d);
  }
  if (rc) return rc;
  std::lock_guard<std::mutex> dl(g_dev_mu[d->dev]);
  cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
  if (ncell == 0) return 0;

  // device staging: [rconst | fix | var | hexit | texit | ierr | stats], 256 B aligned sections
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t n = (size_t)ncell;
  const size_t b_rc = al(n * mi->nreact * 8), b_fx = al(n * mi->nfix * 8), b_vr = al(n * mi->nvar * 8);
  const size_t b_hx = al(n * 8), b_ie = al(n * 4), b_st = al(n * 32);
  const size_t total = b_rc + b_fx + b_vr + 2 * b_hx + b_ie + b_st;
  if (d->d_stage_bytes < total) {
    if (d->d_stage) cudaFree(d->d_stage);
    d->d_stage = nullptr;
    d->d_stage_bytes = 0;
    CK(cudaMalloc(&d->d_stage, total));
    d->d_stage_bytes = total;
  }
  char *p = (char *)d->d_stage;
  double *d_rc = (double *)p; p += b_rc;
  double *d_fx = (double *)p; p += b_fx;
  double *d_vr = (double *)p; p += b_vr;
  double *d_hx = (double *)p; p += b_hx;
  double *d_tx = (double *)p; p += b_hx;
  int32_t *d_ie = (int32_t *)p; p += b_ie;
  int32_t *d_st = (int32_t *)p;

  // Pipeline over chunks of cells: H2D of chunk c+1 and D2H of chunk c-1 overlap the
  // kernel of chunk c (copy streams + events).  Kernels alternate between two slots - the
  // caller's stream with the first lane workspace / cell counter, an internal stream with the
  // second - so that the CTAs of chunk c+1 move in as those of chunk c run out of cells
  // instead of waiting for its slowest cell (a drained grid per chunk cost ~25 % end to end).
  // Host buffers may be pageable (Fortran arrays) or pinned; cudaMemcpyAsync handles
  // both, only pinned ones actually overlap.
  const bool oc = want_onchip(mi, mech, ncell);     // by the size of the whole batch, not of a chunk
  if ((rc = ensure_mech(*d, mech, mi, o ? (o->f32_literals ? 1 : 0) : 1, st, oc))) return rc;
  Handoff ho, *hop = nullptr;                       // long cells: continued on chip after the last chunk (kpp_batch.h)
  if (!oc && handoff_steps(mi, mech) > 0) {
    if ((rc = ensure_mech(*d, mech, mi, o ? (o->f32_literals ? 1 : 0) : 1, st, true))) return rc;
    if ((rc = handoff_begin(*d, ncell, handoff_steps(mi, mech), &ho, st))) return rc;
    hop = &ho;
  }
  const int64_t resident = oc ? (int64_t)d->mech[mech].oc_blocks * mi->oc->slots : (int64_t)d->mech[mech].blocks * KPP_BLOCK;
  int64_t nchunk = ncell / (2 * resident);
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 16) nchunk = 16;
  if (const char *e = getenv("MISTRA_KPP_CHUNKS")) {
    int v = atoi(e);
    if (v >= 1 && v <= 64) nchunk = v;
  }
  const int64_t per = (ncell + nchunk - 1) / nchunk;
  if ((rc = ensure_streams(d))) return rc;
  // the copy streams must not run ahead of work already queued on the caller's stream
  CK(cudaEventRecord(d->ev_start, st));
  CK(cudaStreamWaitEvent(d->s_h2d, d->ev_start, 0));
  CK(cudaStreamWaitEvent(d->s_d2h, d->ev_start, 0));
  CK(cudaStreamWaitEvent(d->s_k2, d->ev_start, 0));
  // a failure in the middle of the pipeline must not leave copies in flight that still read or write the
  // caller's arrays after the call has returned: the four streams are drained before the error is reported
  auto pipeline = [&]() -> int {
  for (int64_t c = 0, off = 0; off < ncell; ++c, off += per) {
    const int slot = (int)(c & 1);
    cudaStream_t ks = slot ? d->s_k2 : st;
    const size_t m = (size_t)((ncell - off) < per ? (ncell - off) : per), o0 = (size_t)off;
    CK(cudaMemcpyAsync(d_rc + o0 * mi->nreact, rconst + o0 * mi->nreact, m * mi->nreact * 8,
                       cudaMemcpyHostToDevice, d->s_h2d));
    CK(cudaMemcpyAsync(d_fx + o0 * mi->nfix, fix + o0 * mi->nfix, m * mi->nfix * 8,
                       cudaMemcpyHostToDevice, d->s_h2d));
    CK(cudaMemcpyAsync(d_vr + o0 * mi->nvar, var + o0 * mi->nvar, m * mi->nvar * 8,
                       cudaMemcpyHostToDevice, d->s_h2d));
    CK(cudaEventRecord(d->ev_h[c], d->s_h2d));
    CK(cudaStreamWaitEvent(ks, d->ev_h[c], 0));
    rc = launch_device(*d, mech, (int64_t)m, d_rc + o0 * mi->nreact, d_fx + o0 * mi->nfix,
                       d_vr + o0 * mi->nvar, t0, t1, o, ierr ? d_ie + o0 : nullptr,
                       stats ? d_st + o0 * 8 : nullptr, hexit ? d_hx + o0 : nullptr,
                       texit ? d_tx + o0 : nullptr, ks, slot, oc, hop, off);
    if (rc) return rc;
    CK(cudaEventRecord(d->ev_k[c], ks));
  }
  // results: queued after every input copy and kernel, so that a blocking copy into pageable
  // host memory (Fortran arrays) stalls this thread only, not the chunks behind it
  for (int64_t c = 0, off = 0; off < ncell; ++c, off += per) {
    const size_t m = (size_t)((ncell - off) < per ? (ncell - off) : per), o0 = (size_t)off;
    CK(cudaStreamWaitEvent(d->s_d2h, d->ev_k[c], 0));
    CK(cudaMemcpyAsync(var + o0 * mi->nvar, d_vr + o0 * mi->nvar, m * mi->nvar * 8,
                       cudaMemcpyDeviceToHost, d->s_d2h));
    if (ierr) CK(cudaMemcpyAsync(ierr + o0, d_ie + o0, m * 4, cudaMemcpyDeviceToHost, d->s_d2h));
    if (stats) CK(cudaMemcpyAsync(stats + o0 * 8, d_st + o0 * 8, m * 32, cudaMemcpyDeviceToHost, d->s_d2h));
    if (hexit) CK(cudaMemcpyAsync(hexit + o0, d_hx + o0, m * 8, cudaMemcpyDeviceToHost, d->s_d2h));
    if (texit) CK(cudaMemcpyAsync(texit + o0, d_tx + o0, m * 8, cudaMemcpyDeviceToHost, d->s_d2h));
  }
  CK(cudaEventRecord(d->ev_join, d->s_k2));     // the caller's stream ends after both slots
  CK(cudaStreamWaitEvent(st, d->ev_join, 0));
  if (hop && (rc = launch_handoff_pass(*d, mech, ncell, d_rc, d_fx, d_vr, t0, t1, o, ierr ? d_ie : nullptr,
                                       stats ? d_st : nullptr, hexit ? d_hx : nullptr, texit ? d_tx : nullptr, st, ho)))
    return rc;
  CK(cudaStreamSynchronize(d->s_d2h));
  CK(cudaStreamSynchronize(st));
  if (hop) return handoff_fixup_host(*d, mi, ncell, ho, d_vr, d_ie, d_st, d_hx, d_tx, var, ierr, stats, hexit, texit, st);
  return 0;
  };
  rc = pipeline();
  if (rc) {
    const std::string msg = g_err;
    cudaStreamSynchronize(d->s_h2d);
    cudaStreamSynchronize(d->s_k2);
    cudaStreamSynchronize(d->s_d2h);
    cudaStreamSynchronize(st);
    g_err = msg;
  }
  return rc;
}


int mistra_kpp_integrate_rates(int mech, int64_t ncell, const mistra_rate_inputs_compact *rates, const double *fix,
                               double *var, double t0, double t1, const mistra_kpp_opts *o, int32_t *ierr,
                               int32_t *stats, double *hexit, double *texit, int64_t *h2d_bytes, void *stream)
{
  const KppMechInfo *mi = mech_info(mech);
  if (!mi) return fail(MISTRA_KPP_EINVAL, "unknown mechanism id");
  if (ncell < 0) return fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (!rates) return fail(MISTRA_KPP_EINVAL, "null rates");
  if (ncell > 0 && (!fix || !var || !rates->cb1 || !rates->scal || !rates->ph_rat))
    return fail(MISTRA_KPP_EINVAL, "null fix / var / cb1 / scal / ph_rat");
  {
    KppBatch tmp;
    int rc = decode_opts(o, t0, t1, &tmp);
    if (rc) return rc;
  }
  const int nspec = mi->nvar + mi->nfix, nkc = (mech == MISTRA_KPP_TOT) ? 4 : 2;
  struct L { const mistra_rate_list *l; int nk; } lists[6] = {{&rates->yhenry, 1}, {&rates->yxkmt, nkc}, {&rates->ykef, nkc},
                                                              {&rates->ykeb, nkc}, {&rates->yxkmtd, 2}, {&rates->yxeq, 1}};
  for (auto &q : lists) {
    if (q.l->n < 0 || q.l->n > nspec) return fail(MISTRA_KPP_EINVAL, "rate list: bad species count");
    if (q.l->n > 0 && (!q.l->idx || !q.l->val)) return fail(MISTRA_KPP_EINVAL, "rate list: null idx / val");
    for (int j = 0; j < q.l->n; ++j)
      if (q.l->idx[j] < 0 || q.l->idx[j] >= nspec) return fail(MISTRA_KPP_EINVAL, "rate list: species index out of range");
  }
  if (h2d_bytes) *h2d_bytes = 0;
  DeviceState *d;
  int rc;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    rc = get_device(&d);
  }
  if (rc) return rc;
  std::lock_guard<std::mutex> dl(g_dev_mu[d->dev]);
  cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
  if (ncell == 0) return 0;
  const bool oc = want_onchip(mi, mech, ncell);
  if ((rc = ensure_mech(*d, mech, mi, o ? (o->f32_literals ? 1 : 0) : 1, st, oc))) return rc;
  if ((rc = ensure_streams(d))) return rc;
  Handoff ho, *hop = nullptr;                       // long cells: continued on chip after the last chunk (kpp_batch.h)
  if (!oc && handoff_steps(mi, mech) > 0) {
    if ((rc = ensure_mech(*d, mech, mi, o ? (o->f32_literals ? 1 : 0) : 1, st, true))) return rc;
    if ((rc = handoff_begin(*d, ncell, handoff_steps(mi, mech), &ho, st))) return rc;
    hop = &ho;
  }

  // chunks: as mistra_kpp_integrate, but at most 65536 cells each (the expanded rate arrays are per chunk)
  const int64_t resident = oc ? (int64_t)d->mech[mech].oc_blocks * mi->oc->slots : (int64_t)d->mech[mech].blocks * KPP_BLOCK;
  int64_t nchunk = ncell / (2 * resident);
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 16) nchunk = 16;
  if ((ncell + nchunk - 1) / nchunk > 65536) nchunk = (ncell + 65535) / 65536;
  if (nchunk > 64) return fail(MISTRA_KPP_EINVAL, "batch too large for one call (more than 64 x 65536 cells)");
  const int64_t per = (ncell + nchunk - 1) / nchunk;

  // ---- device staging of the whole batch: fix | var | diagnostics | scalars | compact lists | index lists
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t n = (size_t)ncell;
  const size_t b_fx = al(n * mi->nfix * 8), b_vr = al(n * mi->nvar * 8), b_hx = al(n * 8), b_ie = al(n * 4), b_st = al(n * 32);
  const size_t b_cb = al(n * 4 * 8), b_sc = al(n * 13 * 8), b_ph = al(n * 47 * 8), b_cw = al(n * nkc * 8), b_cd = al(n * 2 * 8);
  size_t b_val[6], b_idx[6], tot_lists = 0;
  for (int q = 0; q < 6; ++q) {
    b_val[q] = al(n * lists[q].nk * (size_t)lists[q].l->n * 8);
    b_idx[q] = al((size_t)lists[q].l->n * 4 + 4);
    tot_lists += b_val[q] + b_idx[q];
  }
  const size_t total = b_fx + b_vr + 2 * b_hx + b_ie + b_st + b_cb + b_sc + b_ph + b_cw + b_cd + tot_lists;
  if (d->d_stage_bytes < total) {
    CK(cudaDeviceSynchronize());
    if (d->d_stage) cudaFree(d->d_stage);
    d->d_stage = nullptr;
    d->d_stage_bytes = 0;
    CK(cudaMalloc(&d->d_stage, total));
    d->d_stage_bytes = total;
  }
  char *p = (char *)d->d_stage;
  double *d_fx = (double *)p; p += b_fx;
  double *d_vr = (double *)p; p += b_vr;
  double *d_hx = (double *)p; p += b_hx;
  double *d_tx = (double *)p; p += b_hx;
  int32_t *d_ie = (int32_t *)p; p += b_ie;
  int32_t *d_st = (int32_t *)p; p += b_st;
  double *d_cb = (double *)p; p += b_cb;
  double *d_sc = (double *)p; p += b_sc;
  double *d_ph = (double *)p; p += b_ph;
  double *d_cw = (double *)p; p += b_cw;
  double *d_cd = (double *)p; p += b_cd;
  double *d_val[6];
  int *d_idx[6];
  for (int q = 0; q < 6; ++q) { d_val[q] = (double *)p; p += b_val[q]; d_idx[q] = (int *)p; p += b_idx[q]; }

  // ---- per-slot scratch: RCONST of a chunk and the NSPEC-indexed arrays (zero outside the lists, set once per call)
  // (with the hand-off the RCONST of every chunk must survive until the last pass: one block for the whole batch,
  // laid out behind the two slots, device memory only)
  const size_t c_rc = hop ? 0 : al((size_t)per * mi->nreact * 8);
  const size_t c_rc_all = hop ? al(n * mi->nreact * 8) : 0;
  const int nk_full[7] = {1, nkc, nkc, nkc, 2, 1, 1};                 // yhenry yxkmt ykef ykeb yxkmtd yxeq | conc
  size_t c_full[7], c_slot = c_rc;
  for (int q = 0; q < 7; ++q) { c_full[q] = al((size_t)per * nk_full[q] * nspec * 8); c_slot += c_full[q]; }
  if (d->d_rates_bytes < 2 * c_slot + c_rc_all) {
    CK(cudaDeviceSynchronize());
    if (d->d_rates) cudaFree(d->d_rates);
    d->d_rates = nullptr;
    d->d_rates_bytes = 0;
    CK(cudaMalloc(&d->d_rates, 2 * c_slot + c_rc_all));
    d->d_rates_bytes = 2 * c_slot + c_rc_all;
  }
  double *s_rc[2], *s_full[2][7];
  double *d_rc_all = hop ? (double *)((char *)d->d_rates + 2 * c_slot) : nullptr;
  for (int sl = 0; sl < 2; ++sl) {
    char *q = (char *)d->d_rates + sl * c_slot;
    s_rc[sl] = (double *)q; q += c_rc;
    for (int k = 0; k < 7; ++k) { s_full[sl][k] = (double *)q; q += c_full[k]; }
  }
  int64_t moved = 0;
  auto pipeline = [&]() -> int {
    CK(cudaEventRecord(d->ev_start, st));
    CK(cudaStreamWaitEvent(d->s_h2d, d->ev_start, 0));
    CK(cudaStreamWaitEvent(d->s_d2h, d->ev_start, 0));
    CK(cudaStreamWaitEvent(d->s_k2, d->ev_start, 0));
    for (int sl = 0; sl < 2; ++sl)          // entries outside the index lists stay zero for the whole call
      CK(cudaMemsetAsync(s_full[sl][0], 0, c_slot - c_rc, sl ? d->s_k2 : st));
    for (int q = 0; q < 6; ++q)
      if (lists[q].l->n > 0) {
        CK(cudaMemcpyAsync(d_idx[q], lists[q].l->idx, (size_t)lists[q].l->n * 4, cudaMemcpyHostToDevice, d->s_h2d));
        moved += (int64_t)lists[q].l->n * 4;
      }
    for (int64_t c = 0, off = 0; off < ncell; ++c, off += per) {
      const int slot = (int)(c & 1);
      cudaStream_t ks = slot ? d->s_k2 : st;
      const size_t m = (size_t)((ncell - off) < per ? (ncell - off) : per), o0 = (size_t)off;
      auto up = [&](double *dst, const double *src, size_t w) -> cudaError_t {
        moved += (int64_t)(m * w * 8);
        return cudaMemcpyAsync(dst + o0 * w, src + o0 * w, m * w * 8, cudaMemcpyHostToDevice, d->s_h2d);
      };
      CK(up(d_fx, fix, mi->nfix));
      CK(up(d_vr, var, mi->nvar));
      CK(up(d_cb, rates->cb1, 4));
      CK(up(d_sc, rates->scal, 13));
      CK(up(d_ph, rates->ph_rat, 47));
      if (rates->ycw) CK(up(d_cw, rates->ycw, nkc));
      if (rates->ycwd) CK(up(d_cd, rates->ycwd, 2));
      for (int q = 0; q < 6; ++q)
        if (lists[q].l->n > 0) CK(up(d_val[q], lists[q].l->val, (size_t)lists[q].nk * lists[q].l->n));
      CK(cudaEventRecord(d->ev_h[c], d->s_h2d));
      CK(cudaStreamWaitEvent(ks, d->ev_h[c], 0));
      // expand -> Update_RCONST_x -> INTEGRATE_x on the chunk's stream
      for (int q = 0; q < 6; ++q) {
        const int nl = lists[q].l->n;
        if (nl == 0) continue;
        const long long nrow = (long long)m * lists[q].nk, tot = nrow * nl;
        rates_expand_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ks>>>(s_full[slot][q], d_val[q] + o0 * lists[q].nk * nl,
                                                                             d_idx[q], nrow, nl, nspec);
      }
      {
        const long long tot = (long long)m * nspec;
        rates_conc_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ks>>>(s_full[slot][6], d_vr + o0 * mi->nvar, d_fx + o0 * mi->nfix,
                                                                         (long long)m, mi->nvar, mi->nfix);
      }
      CK(cudaGetLastError());
      mistra_rate_inputs in;
      memset(&in, 0, sizeof in);
      in.ncell = (int64_t)m;
      in.cb1 = d_cb + o0 * 4; in.scal = d_sc + o0 * 13; in.ph_rat = d_ph + o0 * 47; in.conc = s_full[slot][6];
      in.yhenry = s_full[slot][0]; in.yxkmt = s_full[slot][1]; in.ykef = s_full[slot][2]; in.ykeb = s_full[slot][3];
      in.yxkmtd = s_full[slot][4]; in.yxeq = s_full[slot][5];
      in.ycw = rates->ycw ? d_cw + o0 * nkc : nullptr;
      in.ycwd = rates->ycwd ? d_cd + o0 * 2 : nullptr;
      in.f32_literals = rates->f32_literals;
      double *rc_chunk = hop ? d_rc_all + o0 * mi->nreact : s_rc[slot];
      if (int r2 = mistra_rconst_update_device(mech, &in, rc_chunk, ks)) return r2;
      int r3 = launch_device(*d, mech, (int64_t)m, rc_chunk, d_fx + o0 * mi->nfix, d_vr + o0 * mi->nvar, t0, t1, o,
                             ierr ? d_ie + o0 : nullptr, stats ? d_st + o0 * 8 : nullptr, hexit ? d_hx + o0 : nullptr,
                             texit ? d_tx + o0 : nullptr, ks, slot, oc, hop, off);
      if (r3) return r3;
      CK(cudaEventRecord(d->ev_k[c], ks));
    }
    for (int64_t c = 0, off = 0; off < ncell; ++c, off += per) {
      const size_t m = (size_t)((ncell - off) < per ? (ncell - off) : per), o0 = (size_t)off;
      CK(cudaStreamWaitEvent(d->s_d2h, d->ev_k[c], 0));
      CK(cudaMemcpyAsync(var + o0 * mi->nvar, d_vr + o0 * mi->nvar, m * mi->nvar * 8, cudaMemcpyDeviceToHost, d->s_d2h));
      if (ierr) CK(cudaMemcpyAsync(ierr + o0, d_ie + o0, m * 4, cudaMemcpyDeviceToHost, d->s_d2h));
      if (stats) CK(cudaMemcpyAsync(stats + o0 * 8, d_st + o0 * 8, m * 32, cudaMemcpyDeviceToHost, d->s_d2h));
      if (hexit) CK(cudaMemcpyAsync(hexit + o0, d_hx + o0, m * 8, cudaMemcpyDeviceToHost, d->s_d2h));
      if (texit) CK(cudaMemcpyAsync(texit + o0, d_tx + o0, m * 8, cudaMemcpyDeviceToHost, d->s_d2h));
    }
    CK(cudaEventRecord(d->ev_join, d->s_k2));
    CK(cudaStreamWaitEvent(st, d->ev_join, 0));
    if (hop)
      if (int r4 = launch_handoff_pass(*d, mech, ncell, d_rc_all, d_fx, d_vr, t0, t1, o, ierr ? d_ie : nullptr,
                                       stats ? d_st : nullptr, hexit ? d_hx : nullptr, texit ? d_tx : nullptr, st, ho))
        return r4;
    CK(cudaStreamSynchronize(d->s_d2h));
    CK(cudaStreamSynchronize(st));
    if (hop) return handoff_fixup_host(*d, mi, ncell, ho, d_vr, d_ie, d_st, d_hx, d_tx, var, ierr, stats, hexit, texit, st);
    return 0;
  };
  rc = pipeline();
  if (rc) {
    const std::string msg = g_err;
    cudaStreamSynchronize(d->s_h2d);
    cudaStreamSynchronize(d->s_k2);
    cudaStreamSynchronize(d->s_d2h);
    cudaStreamSynchronize(st);
    g_err = msg;
  }
  if (h2d_bytes) *h2d_bytes = moved;
  return rc;
}

int64_t mistra_kpp_launch_count(void) { return g_launches.load(); }

// ---- page-locked host memory: what lets the chunk pipeline overlap its copies with the kernels ----------------
int mistra_kpp_host_alloc(void **p, size_t bytes)
{
  if (!p) return fail(MISTRA_KPP_EINVAL, "null pointer");
  *p = nullptr;
  if (bytes == 0) return 0;
  CK(cudaHostAlloc(p, bytes, cudaHostAllocPortable));
  return 0;
}

int mistra_kpp_host_free(void *p)
{
  if (p) CK(cudaFreeHost(p));
  return 0;
}

int mistra_kpp_host_register(void *p, size_t bytes)
{
  if (!p || bytes == 0) return fail(MISTRA_KPP_EINVAL, "null array");
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}

int mistra_kpp_host_unregister(void *p)
{
  if (!p) return fail(MISTRA_KPP_EINVAL, "null array");
  CK(cudaHostUnregister(p));
  return 0;
}

// ---- all GPUs of the box from one process ------------------------------------------------------------------
int mistra_kpp_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  return n;
}

int mistra_kpp_integrate_multi(int mech, int64_t ncell, const double *rconst, const double *fix, double *var,
                               double t0, double t1, const mistra_kpp_opts *o, int32_t *ierr, int32_t *stats,
                               double *hexit, double *texit, int ndev, const int *devices)
{
  const KppMechInfo *mi = mech_info(mech);
  if (!mi) return fail(MISTRA_KPP_EINVAL, "unknown mechanism id");
  if (ncell < 0) return fail(MISTRA_KPP_EINVAL, "ncell < 0");
  if (ncell > 0 && (!rconst || !fix || !var)) return fail(MISTRA_KPP_EINVAL, "null rconst/fix/var");
  {
    KppBatch tmp;
    int rc = decode_opts(o, t0, t1, &tmp);
    if (rc) return rc;
  }
  int visible = 0;
  CK(cudaGetDeviceCount(&visible));
  if (visible < 1) return fail(MISTRA_KPP_ENODEVICE, "no CUDA device");
  if (ndev <= 0) { ndev = visible; devices = nullptr; }
  if (ndev > 64) return fail(MISTRA_KPP_EINVAL, "more than 64 slices");
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; ++i) {
    devs[i] = devices ? devices[i] : i;
    if (devs[i] < 0 || devs[i] >= visible || devs[i] >= kMaxDev) return fail(MISTRA_KPP_EINVAL, "device index out of range");
  }
  int cur = 0;
  CK(cudaGetDevice(&cur));
  // contiguous slices of cells, one host thread per slice (a device listed twice gets two slices, one after the other)
  std::vector<int> rcs(ndev, 0);
  std::vector<std::string> msgs(ndev);
  auto work = [&](int i) {
    const int64_t lo = ncell * i / ndev, hi = ncell * (i + 1) / ndev, m = hi - lo;
    if (m <= 0) return;
    cudaError_t e = cudaSetDevice(devs[i]);
    if (e != cudaSuccess) { rcs[i] = cuda_fail(e, "cudaSetDevice"); msgs[i] = g_err; return; }
    rcs[i] = mistra_kpp_integrate(mech, m, rconst + lo * mi->nreact, fix + lo * mi->nfix, var + lo * mi->nvar, t0, t1, o,
                                  ierr ? ierr + lo : nullptr, stats ? stats + lo * 8 : nullptr,
                                  hexit ? hexit + lo : nullptr, texit ? texit + lo : nullptr, nullptr);
    if (rcs[i]) msgs[i] = g_err;
  };
  std::vector<std::thread> th;
  for (int i = 1; i < ndev; ++i) th.emplace_back(work, i);
  work(0);
  for (auto &t : th) t.join();
  cudaSetDevice(cur);
  for (int i = 0; i < ndev; ++i)
    if (rcs[i]) { g_err = "slice " + std::to_string(i) + " (device " + std::to_string(devs[i]) + "): " + msgs[i]; return rcs[i]; }
  return 0;
}

int mistra_kpp_set_kernel(int mech, int variant)
{
  if (mech < 0 || mech > 2) return fail(MISTRA_KPP_EINVAL, "mech must be 0, 1 or 2");
  if (variant < -1 || variant > 1) return fail(MISTRA_KPP_EINVAL, "variant must be -1 (by batch size), 0 (cell per thread) or 1 (on-chip)");
  if (variant == 1 && !mech_info(mech)->oc) return fail(MISTRA_KPP_EINVAL, "this mechanism has no on-chip kernel");
  std::lock_guard<std::mutex> lk(g_mu);
  g_variant[mech] = variant;
  return 0;
}

int mistra_kpp_set_handoff(int mech, int steps)
{
  if (!mech_info(mech)) return fail(MISTRA_KPP_EINVAL, "unknown mechanism id");
  if (steps < -1) return fail(MISTRA_KPP_EINVAL, "steps: -1 default, 0 off, > 0 step attempts before the hand-off");
  if (steps > 0 && !mech_info(mech)->oc) return fail(MISTRA_KPP_EINVAL, "this mechanism has no on-chip kernel to hand over to");
  std::lock_guard<std::mutex> lk(g_mu);
  g_handoff[mech] = steps;
  return 0;
}

int64_t mistra_kpp_handoff_count(void)
{
  DeviceState *d;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (get_device(&d)) return -1;
  }
  std::lock_guard<std::mutex> dl(g_dev_mu[d->dev]);
  if (!d->handoff) return 0;
  unsigned long long h = 0;
  if (cudaMemcpy(&h, d->handoff, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)h;
}

int mistra_kpp_get_kernel(int mech)
{
  if (mech < 0 || mech > 2) return fail(MISTRA_KPP_EINVAL, "mech must be 0, 1 or 2");
  return g_variant[mech];
}

int mistra_kpp_kernel_for(int mech, int64_t ncell)
{
  if (mech < 0 || mech > 2) return fail(MISTRA_KPP_EINVAL, "mech must be 0, 1 or 2");
  return want_onchip(mech_info(mech), mech, ncell) ? 1 : 0;
}

int64_t mistra_kpp_launch_count_variant(int variant)
{
  return (variant == 0 || variant == 1) ? g_launches_variant[variant].load() : -1;
}

// Internal (not part of include/mistra_kpp.h): hands a device buffer and a flag word to the on-chip kernels;
// only builds with -DOC_DEBUG / -DOC_PHASE_TIMERS look at them.
extern "C" void mistra_kpp_oc_debug(void *dev_buffer, long long flags)
{
  g_oc_aux = dev_buffer;
  g_oc_flags = flags;
}

int mistra_kpp_finalize(void)
{
  std::lock_guard<std::mutex> lk(g_mu);
  int cur = -1;
  cudaGetDevice(&cur);
  for (int i = 0; i < kMaxDev; ++i) {
    std::lock_guard<std::mutex> dl(g_dev_mu[i]);
    DeviceState &d = g_dev[i];
    if (!d.init) continue;
    cudaSetDevice(i);
    cudaDeviceSynchronize();
    for (auto &m : d.mech) {
      if (m.ws) cudaFree(m.ws);
      if (m.ws2) cudaFree(m.ws2);
      if (m.oc_tab) cudaFree(m.oc_tab);
      if (m.oc_ws) cudaFree(m.oc_ws);
      if (m.oc_ws2) cudaFree(m.oc_ws2);
      m = MechState();
    }
    if (d.counter) cudaFree(d.counter);
    if (d.handoff) cudaFree(d.handoff);
    d.handoff = nullptr; d.handoff_bytes = 0;
    if (d.fix_dev) cudaFree(d.fix_dev);
    if (d.fix_host) cudaFreeHost(d.fix_host);
    d.fix_dev = d.fix_host = nullptr; d.fix_bytes = 0;
    for (auto &e : d.ev_slot) if (e) cudaEventDestroy(e);
    if (d.d_stage) cudaFree(d.d_stage);
    if (d.d_rates) cudaFree(d.d_rates);
    if (d.s_h2d) {
      cudaStreamDestroy(d.s_h2d);
      cudaStreamDestroy(d.s_d2h);
      cudaStreamDestroy(d.s_k2);
      cudaEventDestroy(d.ev_join);
      for (auto &e : d.ev_h) cudaEventDestroy(e);
      for (auto &e : d.ev_k) cudaEventDestroy(e);
      cudaEventDestroy(d.ev_start);
    }
    if (d.stream) cudaStreamDestroy(d.stream);
    d = DeviceState();
  }
  if (cur >= 0) cudaSetDevice(cur);
  return 0;
}

const char *mistra_kpp_last_error(void) { return g_err.c_str(); }

}  // extern "C"
