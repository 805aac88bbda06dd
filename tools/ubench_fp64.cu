// Micro-benchmarks that size the on-chip Ros3 design (tools only, not product code):
// FP64 dependent latency / throughput, shuffle, shared-memory and barrier costs on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ---- single-warp dependent chains, timed with clock64 ---------------------------------
template <int OP>
__global__ void lat_kernel(double *out, long long *cyc, int iters, double a, double b)
{
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = a + i * 1e-9;
  __syncthreads();
  double x = threadIdx.x * 1e-3 + 1.0;
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (OP == 0) x = fma(x, a, b);
      else if (OP == 1) x = x * a;
      else if (OP == 2) x = x + b;
      else if (OP == 3) x = b / x;                        // IEEE division
      else if (OP == 4) x = __shfl_sync(0xffffffffu, x, (u * 7 + 3) & 31);
      else if (OP == 5) { x = sm[idx]; idx = (int)(x) & 1023; idx = (idx + threadIdx.x) & 1023; } // LDS dependent (with cvt)
      else if (OP == 6) { x = fma(x, a, sm[(u * 33) & 1023]); }   // DFMA with smem broadcast operand
      else if (OP == 7) { x = __shfl_sync(0xffffffffu, x, u & 31); x = fma(x, a, b); }  // shuffle + fma chain (column sweep step)
      else if (OP == 8) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + b; __syncwarp(); }  // STS/LDS round trip
      else if (OP == 9) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + b; }
      else if (OP == 10) { x = 1.0 / x + b; }
      else if (OP == 11) { x = sqrt(x) + b; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (x == 12345.678) out[0] = x;
}

// ---- barrier latency: nthreads per block, loops of bar.sync ---------------------------
__global__ void bar_kernel(long long *cyc, int iters)
{
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// named barrier over a subset (64 threads) while other warps idle
__global__ void namedbar_kernel(long long *cyc, int iters, int nthr)
{
  const int grp = threadIdx.x / nthr;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthr));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ---- throughput: ILP independent DFMA chains, W warps per block, 1 block per SM --------
template <int ILP, int SRC>
__global__ void thr_kernel(double *out, long long *cyc, int iters, double a, double b)
{
  __shared__ double sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = a + i * 1e-9;
  __syncthreads();
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (SRC == 0) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
      } else if (SRC == 1) {   // one smem broadcast operand per FMA (LDS.64)
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, sm[(u * ILP + j + (i & 7) * 64) & 2047]);
      } else if (SRC == 2) {   // LDS.128 broadcast: 2 operands per load
#pragma unroll
        for (int j = 0; j < ILP; j += 2) {
          const double2 v = *reinterpret_cast<const double2 *>(&sm[((u * ILP + j) + (i & 7) * 64) & 2046]);
          x[j] = fma(x[j], a, v.x);
          x[j + 1] = fma(x[j + 1], a, v.y);
        }
      } else if (SRC == 3) {   // per-lane (conflict-free) smem operand per FMA
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, sm[((u * ILP + j) * 32 + (threadIdx.x & 31) + (i & 1) * 1024) & 2047]);
      } else if (SRC == 4) {   // per-lane smem read-modify-write: sm[...] -= a * x  (LU update out of smem)
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
          const int id = ((u * ILP + j) * 32 + (threadIdx.x & 31)) & 2047;
          sm[id] = fma(x[j], a, sm[id]);
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 12345.678) out[0] = s + sm[5];
}

template <int OP>
static void run_lat(const char *name, double *d_out, long long *d_cyc)
{
  const int iters = 2000;
  lat_kernel<OP><<<1, 32>>>(d_out, d_cyc, iters, 0.9999999, 1e-9);
  CK(cudaDeviceSynchronize());
  lat_kernel<OP><<<1, 32>>>(d_out, d_cyc, iters, 0.9999999, 1e-9);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
  printf("LAT %-28s %8.2f cycles/op\n", name, (double)c / (iters * 16.0));
}

template <int ILP, int SRC>
static void run_thr(const char *name, double *d_out, long long *d_cyc, int warps)
{
  const int iters = 2000;
  thr_kernel<ILP, SRC><<<1, warps * 32>>>(d_out, d_cyc, iters, 0.9999999, 1e-9);
  CK(cudaDeviceSynchronize());
  thr_kernel<ILP, SRC><<<1, warps * 32>>>(d_out, d_cyc, iters, 0.9999999, 1e-9);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
  const double fma_per_clk = (double)iters * 4 * ILP * warps * 32 / (double)c;
  printf("THR %-22s warps=%2d ILP=%2d : %7.2f FMA/clk/SM  (%.2f cycles per warp-instr per SMSP)\n", name, warps, ILP,
         fma_per_clk, (double)c / ((double)iters * 4 * ILP * warps / 4.0));
}

int main()
{
  double *d_out;
  long long *d_cyc;
  CK(cudaMalloc(&d_out, 64));
  CK(cudaMalloc(&d_cyc, 8 * 1024));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run_lat<0>("dfma dependent", d_out, d_cyc);
  run_lat<1>("dmul dependent", d_out, d_cyc);
  run_lat<2>("dadd dependent", d_out, d_cyc);
  run_lat<3>("ddiv (b/x) dependent", d_out, d_cyc);
  run_lat<4>("shfl f64 dependent", d_out, d_cyc);
  run_lat<5>("lds+cvt dependent", d_out, d_cyc);
  run_lat<6>("dfma + lds bcast operand", d_out, d_cyc);
  run_lat<7>("shfl + dfma", d_out, d_cyc);
  run_lat<8>("sts/syncwarp/lds/dadd", d_out, d_cyc);
  run_lat<9>("rcp.approx.f64 + dadd", d_out, d_cyc);
  run_lat<10>("1.0/x + dadd", d_out, d_cyc);
  run_lat<11>("sqrt + dadd", d_out, d_cyc);
  for (int nthr : {32, 64, 128, 256, 512}) {
    bar_kernel<<<1, nthr>>>(d_cyc, 1000);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("BAR __syncthreads %3d threads: %6.2f cycles\n", nthr, (double)c / 16000.0);
  }
  for (int nthr : {64, 128}) {
    namedbar_kernel<<<1, 256>>>(d_cyc, 1000, nthr);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("BAR named, groups of %3d in a 256-thread CTA: %6.2f cycles\n", nthr, (double)c / 16000.0);
  }
  for (int warps : {4, 8, 16, 32}) {
    run_thr<1, 0>("dfma reg", d_out, d_cyc, warps);
    run_thr<2, 0>("dfma reg", d_out, d_cyc, warps);
    run_thr<4, 0>("dfma reg", d_out, d_cyc, warps);
    run_thr<8, 0>("dfma reg", d_out, d_cyc, warps);
    run_thr<16, 0>("dfma reg", d_out, d_cyc, warps);
  }
  for (int warps : {4, 8, 16}) {
    run_thr<8, 1>("dfma+lds64 bcast", d_out, d_cyc, warps);
    run_thr<16, 1>("dfma+lds64 bcast", d_out, d_cyc, warps);
    run_thr<8, 2>("dfma+lds128 bcast", d_out, d_cyc, warps);
    run_thr<16, 2>("dfma+lds128 bcast", d_out, d_cyc, warps);
    run_thr<8, 3>("dfma+lds64 per-lane", d_out, d_cyc, warps);
    run_thr<16, 3>("dfma+lds64 per-lane", d_out, d_cyc, warps);
    run_thr<8, 4>("smem rmw per-lane", d_out, d_cyc, warps);
    run_thr<16, 4>("smem rmw per-lane", d_out, d_cyc, warps);
  }
  return 0;
}
