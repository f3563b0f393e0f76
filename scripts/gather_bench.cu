// Micro-benchmark behind DESIGN.md §6 / §9: how fast can an SM gather 16-byte vector entries (the x[col] loads of the fused
// complex SpMV) and feed them to 4 DFMAs each, as a function of
//   * warps per SM (occupancy), * gathers in flight per lane (U), * the size of the gathered array (L1 / L2 residency),
//   * the locality of the columns (a band around the row, or anywhere), * the lane layout: "member" = every lane its own
//     column (up to 32 lines of 128 B per warp-load) or "interleaved" = 8 lanes share a column and read the 8 consecutive
//     entries of its 128-byte line (4 lines per warp-load).
// Columns come from a hash in registers (no index loads), so the kernel isolates the gather + FMA part of a pass.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/gather_bench.cu -o gpurun_out/gather_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ double2 ldg16(const double2* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// rows: number of rows of one vector; nvec vectors back to back (member layout) or nvec/8 groups of [row][8] (interleaved)
template <int U, bool INTERLEAVED>
__global__ void __launch_bounds__(1024, 2) k_gather(const double2* __restrict__ x, int rows, int nvec, int band, int rounds, double* sink) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  double ar = 0.0, ai = 0.0;
  for (int r = 0; r < rounds; ++r) {
    // a "slice": 32 consecutive rows (member layout) or 4 (interleaved) of one vector / group
    const uint32_t item = (uint32_t)(r * nw + gw);
    const int vec = INTERLEAVED ? (int)(item % (uint32_t)(nvec / 8)) : (int)(item % (uint32_t)nvec);
    const int row0 = (int)(hash32(item * 2654435761u) % (uint32_t)(rows - 64));
    const int row = INTERLEAVED ? row0 + (lane >> 3) : row0 + lane;
    double2 xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t h = hash32((uint32_t)row * 40503u + (uint32_t)u * 9973u + item);
      int col = band > 0 ? row + (int)(h % (uint32_t)(2 * band)) - band : (int)(h % (uint32_t)rows);
      col = col < 0 ? 0 : (col >= rows ? rows - 1 : col);
      const double2* p = INTERLEAVED ? x + ((size_t)vec * rows + col) * 8 + (lane & 7) : x + (size_t)vec * rows + col;
      xv[u] = ldg16(p);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double pa = 1.0 + 1e-9 * u, pb = 0.5;
      ar = fma(pa, xv[u].x, ar);
      ar = fma(-pb, xv[u].y, ar);
      ai = fma(pa, xv[u].y, ai);
      ai = fma(pb, xv[u].x, ai);
    }
  }
  if (ar + ai == 12345.678) sink[0] = ar;   // keep the arithmetic
}

template <int U, bool IL>
static void run(const double2* x, int rows, int nvec, int band, int threads, int blocks_per_sm, int sms, double* sink, double clk_ghz) {
  const int rounds = 64;
  const int grid = sms * blocks_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) k_gather<U, IL><<<grid, threads>>>(x, rows, nvec, band, rounds, sink);
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int w = 0; w < reps; ++w) k_gather<U, IL><<<grid, threads>>>(x, rows, nvec, band, rounds, sink);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  const double lane_gathers = (double)reps * rounds * U * (double)grid * threads;
  const double per_s = lane_gathers / (ms * 1e-3);
  printf("%-11s U=%d warps/SM=%2d vectors=%2d (%6.1f MB) band=%6d : %7.1f G lane-gathers/s = %5.2f per clk per SM | 8.4 M take %6.1f us\n",
         IL ? "interleaved" : "member", U, threads / 32 * blocks_per_sm, nvec, nvec * (double)rows * 16 / 1e6, band, per_s * 1e-9,
         per_s / (sms * clk_ghz * 1e9), 8.4e6 / per_s * 1e6);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double clk = clk_khz * 1e-6;
  printf("%s, %d SMs, %.2f GHz; 8.4 M = nonzero-member products of one 16-member pass phase on the 46 k-vertex HARDI mesh\n", prop.name, sms, clk);
  const int rows = 46021;
  const int maxvec = 32;
  double2* x;
  double* sink;
  CK(cudaMalloc(&x, sizeof(double2) * (size_t)rows * maxvec));
  CK(cudaMemset(x, 0, sizeof(double2) * (size_t)rows * maxvec));
  CK(cudaMalloc(&sink, 8));
  const int bands[2] = {1024, 0};
  for (int b = 0; b < 2; ++b) {
    const int band = bands[b];
    for (int nvec : {1, 16}) {
      printf("-- columns %s, %d vector(s)\n", band ? "within +-1024 rows of the row" : "anywhere in the vector", nvec);
      // occupancy sweep at U = 4
      run<4, false>(x, rows, nvec, band, 256, 1, sms, sink, clk);
      run<4, false>(x, rows, nvec, band, 512, 1, sms, sink, clk);
      run<4, false>(x, rows, nvec, band, 1024, 1, sms, sink, clk);
      run<4, false>(x, rows, nvec, band, 1024, 2, sms, sink, clk);
      // gathers in flight at 32 warps per SM
      run<2, false>(x, rows, nvec, band, 1024, 1, sms, sink, clk);
      run<8, false>(x, rows, nvec, band, 1024, 1, sms, sink, clk);
      if (nvec >= 8) {
        run<4, true>(x, rows, nvec, band, 1024, 1, sms, sink, clk);
        run<8, true>(x, rows, nvec, band, 1024, 1, sms, sink, clk);
        run<8, true>(x, rows, nvec, band, 1024, 2, sms, sink, clk);
      }
    }
  }
  return 0;
}
