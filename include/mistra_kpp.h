/* mistra_kpp.h - C ABI of the B200 KPP chemistry path (libmistra_kpp.so).
 *
 * Drop-in boundary for the reference's stiff-chemistry hot path
 * (SURVEY.md §8b).  The reference has no plugin/FFI interface; the path sits
 * behind Fortran external procedures and COMMON blocks, so the boundary is cut at
 * the two places a maintainer can splice a C call in:
 *
 *   B2 (batched, the throughput path): replaces the per-layer body
 *      "Update_RCONST_x ; INTEGRATE_x(tkpp, tkpp+dt_ch)" of the k-loop of
 *      kpp_driver (/root/reference/src/kpp.f90:4310-4470, dispatch 4451-4468;
 *      call sites gas.f:172-173, aer.f:216-217, tot.f:603-604).  The host keeps
 *      gathering C/FIX and computing RCONST per cell exactly as today, stacks
 *      them for all layers of a mechanism, and makes ONE call.
 *
 *   B1 (per-cell link-level shim, see mistra_kpp_f77.h): replaces
 *      INTEGRATE_g/_a/_t(TIN,TOUT) (gas.f:710, aer.f:1408, tot.f:2812).
 *
 * Species order = ind_* of *_Parameters.h, reaction order = RCONST(1:NREACT) of
 * Update_RCONST_x, sparse order = LU_* of *_Sparse.h (all reproduced from the
 * reference's generated sources; check with mistra_kpp_spc_name).
 *
 * All functions return 0 on success or a negative MISTRA_KPP_E* code; the text
 * of the last error is available from mistra_kpp_last_error().  There is no CPU
 * fallback: without a CUDA device every compute entry fails with
 * MISTRA_KPP_ENODEVICE.
 */
#ifndef MISTRA_KPP_H
#define MISTRA_KPP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MISTRA_KPP_GAS = 0, MISTRA_KPP_AER = 1, MISTRA_KPP_TOT = 2 };

enum {
  MISTRA_KPP_OK = 0,
  MISTRA_KPP_EINVAL = -1,    /* bad mechanism id / null pointer / negative size        */
  MISTRA_KPP_EOPTS = -2,     /* option rejected; same tests as Rosenbrock_x gas.f:950-1051 */
  MISTRA_KPP_ENODEVICE = -3, /* no usable CUDA device                                   */
  MISTRA_KPP_ECUDA = -4,     /* CUDA runtime error (see mistra_kpp_last_error)          */
  MISTRA_KPP_ENOMEM = -5
};

/* Integrator options.  Field meaning = RPAR/IPAR of Rosenbrock_x (gas.f:786-870):
 * a zero selects the reference default.  mistra_kpp_default_opts() fills the
 * values INTEGRATE_x hard-codes (gas.f:739-746): Ros3, non-autonomous, scalar
 * tolerances RTOL=1e-3 ATOL=1e-25, Hstart=1e-3; the rest default inside
 * Rosenbrock_x to Hmin=0, Hmax=|t1-t0|, FacMin=.2, FacMax=6, FacRej=.1,
 * FacSafe=.9, 100000 steps. */
typedef struct mistra_kpp_opts {
  double rtol, atol;                      /* RTOL(1), ATOL(1)  (IPAR(2)=1: scalar) */
  double hmin, hmax, hstart;              /* RPAR(1..3) */
  double facmin, facmax, facrej, facsafe; /* RPAR(4..7) */
  int32_t max_steps;                      /* IPAR(3) */
  int32_t autonomous;                     /* IPAR(1)!=0; only changes the Nfun statistic */
  int32_t f32_literals;                   /* 1 (default): non-integer stoichiometric literals
                                             are binary32 as under the reference's preferred
                                             compiler flags; 0: binary64 (-r8 build) */
  int32_t reserved;
} mistra_kpp_opts;

void mistra_kpp_default_opts(mistra_kpp_opts *o);

/* Sizes of a mechanism (gas_Parameters.h:26-49 and the aer_/tot_ twins). */
int mistra_kpp_query(int mech, int *nvar, int *nfix, int *nreact, int *lu_nonzero);

/* SPC_NAMES(i+1) of the mechanism (gas.f:6867-6891), i in [0, NVAR+NFIX); NULL if out of
 * range.  mk_interface (utils.f90:84-140) matches species by these names. */
const char *mistra_kpp_spc_name(int mech, int i);

/* Per-cell exit codes written to ierr[] - those of ros_ErrorMsg_x (gas.f:1474-1509):
 *   1 success; -6 more than max_steps steps; -7 step size too small;
 *  -8 matrix repeatedly singular.  As in the reference (gas.f:764-770) a failed cell
 * returns its partially advanced VAR. */

/* B2 - integrate ncell independent cells of one mechanism from t0 to t1.
 *   rconst [ncell][NREACT]  RCONST of each cell, frozen over the step (gas.f:1964-1967)
 *   fix    [ncell][NFIX]    FIX of each cell
 *   var    [ncell][NVAR]    VAR of each cell, in/out
 *   ierr   [ncell]          or NULL
 *   stats  [ncell][8]       Nfun,Njac,Nstp,Nacc,Nrej,Ndec,Nsol,Nsng (gas.f:913-915) or NULL
 *   hexit  [ncell]          last step size = STEPMIN on return of INTEGRATE_x (gas.f:770) or NULL
 *   texit  [ncell]          time reached = TIN on return of INTEGRATE_x (gas.f:769) or NULL
 * Row-major C arrays = Fortran arrays (NREACT,ncell) etc.  HOST buffers: the call copies them chunk by chunk
 * (cudaMemcpyAsync straight from / to the caller's arrays) to the current CUDA device, runs the chunks on two
 * alternating streams, copies the results back and returns when they are in place.  Pageable arrays work as they
 * are; only page-locked ones (mistra_kpp_host_alloc / mistra_kpp_host_register below) let the copies overlap the
 * kernels.  On a failure in the middle of the pipeline every stream is drained before the error is returned.
 * `stream` is a cudaStream_t (NULL = the library's own stream). */
int mistra_kpp_integrate(int mech, int64_t ncell, const double *rconst, const double *fix,
                         double *var, double t0, double t1, const mistra_kpp_opts *o,
                         int32_t *ierr, int32_t *stats, double *hexit, double *texit,
                         void *stream);

/* Same contract with DEVICE buffers on the current device; asynchronous on
 * `stream` (a cudaStream_t; NULL = the legacy default stream, as in the CUDA
 * runtime), no host synchronisation.  This is the entry the multi-GPU driver
 * and the device-resident benchmark use. */
int mistra_kpp_integrate_device(int mech, int64_t ncell, const double *d_rconst,
                                const double *d_fix, double *d_var, double t0, double t1,
                                const mistra_kpp_opts *o, int32_t *d_ierr, int32_t *d_stats,
                                double *d_hexit, double *d_texit, void *stream);

/* Measured FP64 FMA throughput of the current device in TFLOP/s (an 8-chain DFMA
 * microbenchmark; FMA = 2 flops) - the denominator of the FP64 roofline, since
 * MEASURED_PEAKS.json carries no FP64 figure.  Negative on error. */
double mistra_kpp_fp64_peak_tflops(void);

/* Kernels launched by this library since load (the bench's gpu_launches claim). */
/* The same call over several GPUs of the box from ONE process (what a Fortran kpp_driver needs to reach all 8 GPUs,
 * kpp.f90:4294-4310): the cells are cut into `ndev` contiguous slices, slice i runs on CUDA device devices[i]
 * (devices == NULL: 0..ndev-1; ndev <= 0: every visible device) through mistra_kpp_integrate on its own host
 * thread; no data-path collective, the caller's current device is restored.  Returns the first failing slice's
 * code (message: mistra_kpp_last_error). */
int mistra_kpp_integrate_multi(int mech, int64_t ncell, const double *rconst, const double *fix, double *var,
                               double t0, double t1, const mistra_kpp_opts *o, int32_t *ierr, int32_t *stats,
                               double *hexit, double *texit, int ndev, const int *devices);
int mistra_kpp_device_count(void);

/* Page-locked host memory.  mistra_kpp_integrate takes pageable arrays (Fortran COMMON / module arrays) as they
 * are, but only page-locked ones let its chunk pipeline overlap the copies with the kernels: either allocate the
 * batch arrays here, or register existing arrays once at start-up (e.g. the stacked RCONST / VAR arrays of the
 * patched kpp_driver) and unregister them before they are freed. */
int mistra_kpp_host_alloc(void **p, size_t bytes);
int mistra_kpp_host_free(void *p);
int mistra_kpp_host_register(void *p, size_t bytes);
int mistra_kpp_host_unregister(void *p);

int64_t mistra_kpp_launch_count(void);

/* Kernel variant of a mechanism (a tuning knob, results agree within the parity contract):
 *   0  one cell per thread, per-lane workspace in HBM (csrc/ros3_kernel.inc): needs tens of thousands of cells to fill
 *      the device, the fastest one for large batches;
 *   1  on-chip kernel (csrc/ros3_onchip.inc): one persistent block per SM, cell slots in lockstep, the sparse head of
 *      the LU factors in shared memory, the dense tail in registers, DRAM traffic = compulsory I/O; keeps a few hundred
 *      cells in flight, the faster one for small batches and for the one-cell-per-call box model.  gas and aer only;
 *  -1  (default) chosen per call by the batch size (measured crossover, csrc/kpp_api.cu).
 * mistra_kpp_set_kernel pins a variant for a mechanism (MISTRA_KPP_ONCHIP=0/1 in the environment: for all);
 * mistra_kpp_get_kernel returns the pinned variant or -1, mistra_kpp_kernel_for the variant a batch of ncell cells
 * would run on, mistra_kpp_launch_count_variant the launches of a variant so far.  Returns 0 / a value or a negative
 * MISTRA_KPP_E* code. */
int mistra_kpp_set_kernel(int mech, int variant);
int mistra_kpp_get_kernel(int mech);
int mistra_kpp_kernel_for(int mech, int64_t ncell);
int64_t mistra_kpp_launch_count_variant(int variant);

/* Hand-off of long cells between the two variants.  In the cell-per-thread kernel every Ros3 step of a cell takes
 * milliseconds (tens of thousands of cells share the device), so a cell that needs many steps - rejected steps at a
 * day / night or cloud edge, a transient - keeps its lane long after the rest of the batch has finished and the launch
 * waits for it.  With steps > 0 a cell that has made `steps` step attempts without reaching TOUT is retired at that step
 * boundary and continued, from its (T, H, counters) and VAR, by the on-chip kernel, which runs a step 40 times faster:
 * the same sequence of steps as without the hand-off (the two kernels differ by rounding only, within the parity
 * contract; bit-identical in the strict build).  Which cells are handed over depends on their own step count only, so
 * results do not depend on timing.  steps = -1 (default): 12 for aer when the library chose the kernel by the batch
 * size, off when a variant is pinned, and for gas / tot; 0 = off.  MISTRA_KPP_HANDOFF=<steps> in the environment sets
 * it for aer.  The hand-off pass runs once per call, after the last chunk (mistra_kpp_integrate_device,
 * mistra_kpp_integrate, mistra_kpp_integrate_rates, mistra_kpp_integrate_multi); the host-buffer entries then put the
 * final rows of the handed-over cells in place on the host (mistra_kpp_integrate_rates keeps RCONST of the whole batch
 * in device memory for it).  Returns 0 or a negative MISTRA_KPP_E* code. */
int mistra_kpp_set_handoff(int mech, int steps);
/* Cells the last integration call on the current device handed over (synchronises with the device). */
int64_t mistra_kpp_handoff_count(void);

/* Release what the KPP integrators hold on every device they ran on: lane workspaces, on-chip instruction tables,
 * staging buffers, streams and events.  (The grid caches and scratch buffers of the particle-grid and column
 * operators - mistra_bins.h, mistra_kon.h, ... - are small and live until the process exits.) */
int mistra_kpp_finalize(void);

const char *mistra_kpp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
