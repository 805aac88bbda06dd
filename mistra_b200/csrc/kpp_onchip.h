// Internal: descriptor of a mechanism's on-chip Ros3 kernel (csrc/ros3_onchip.inc, csrc/kpp_onchip_<x>.cu).
#pragma once
#include "kpp_batch.h"

// on-chip kernel of a mechanism (one persistent block per SM with `slots` cells in flight)
struct KppOnchipInfo {
  const void *kernel;
  cudaError_t (*launch)(const KppBatch &, int blocks, cudaStream_t);
  cudaError_t (*set_lit)(const double *host_lit, cudaStream_t);
  const unsigned short *tables;   // host copy of the instruction streams
  size_t table_count;
  const char *const *literals;
  int nlit, threads, smem_bytes, tail, slots;
};

