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
};

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
