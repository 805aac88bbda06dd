// Internal: argument block shared by the per-mechanism kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef KPP_BLOCK
#define KPP_BLOCK 128
#endif
#ifndef KPP_MIN_BLOCKS
#define KPP_MIN_BLOCKS 2
#endif

struct KppBatch {
  // cell data (device, row-major [ncell][...])
  const double *rconst;
  const double *fix;
  double *var;
  int32_t *ierr;    // may be null
  int32_t *stats;   // may be null
  double *hexit;    // may be null
  double *texit;    // may be null
  long long ncell;
  // lane workspace: [resident warps][NSLOT][32 lanes]
  double *ws;
  unsigned long long *counter;  // next unassigned cell, zeroed before launch
  const unsigned short *oc_tab; // on-chip kernels: instruction streams (device copy of the generated tables)
  void *oc_aux;                 // on-chip kernels: spare pointer (diagnostics / experiments)
  long long oc_flags;
  // decoded options (Rosenbrock_x, gas.f:950-1051)
  double t0, t1, rtol, atol, hmin, hmax, hstart, facmin, facmax, facrej, facsafe;
  int max_steps, autonomous;
  // hand-off of long cells (kpp_api.cu: launch_device).  Pass 1 (cell per thread): a cell that has made soft_steps
  // step attempts and is not finished is retired at the step boundary with ierr = KPP_IERR_DEFERRED - VAR, the
  // counters (stats) and (T, H) in cont[cell][2] written out, its index appended to defer_list.  Pass 2 (on-chip
  // kernel, resume = 1) takes the cells list[0 .. *list_count) and continues each from cont / stats: the same
  // sequence of steps as without the hand-off, so a straggler no longer occupies a lane of the slow-per-step kernel.
  int soft_steps, resume;
  long long cell_base;           // index of this launch's first cell in the arrays of the whole call (defer_list holds those)
  long long *defer_list;
  unsigned long long *defer_count;
  double *cont;
  const long long *list;
  const unsigned long long *list_count;
};
#define KPP_IERR_DEFERRED 2

struct KppOnchipInfo;   // csrc/kpp_onchip.h

struct KppMechInfo {
  int nvar, nfix, nreact, lu_nonzero, nslot, ncoef;
  const char *const *coef_literals;
  const void *kernel;  // for occupancy queries
  cudaError_t (*launch)(const KppBatch &, int blocks, cudaStream_t);
  cudaError_t (*set_coef)(const double *host_coef, cudaStream_t);
  const KppOnchipInfo *oc;   // nullptr: this mechanism has no on-chip kernel
};

const KppOnchipInfo *kpp_onchip_info_g();
const KppOnchipInfo *kpp_onchip_info_a();
const KppOnchipInfo *kpp_onchip_info_t();
const KppMechInfo *kpp_mech_info_g();
const KppMechInfo *kpp_mech_info_a();
const KppMechInfo *kpp_mech_info_t();
