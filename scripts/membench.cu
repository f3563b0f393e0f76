// Micro-benchmark (GPU box only): how fast can ~150 MB be streamed once on this B200, cold and warm,
// as a function of loads in flight per thread?  Calibrates the SpMV target and the L2-flush method.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e));return 1;}}while(0)

template <int U>
__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ p, size_t n, double* out) {
  double acc = 0;
  size_t stride = (size_t)gridDim.x * 256 * U;
  for (size_t i = (size_t)blockIdx.x * 256 * U + threadIdx.x; i < n; i += stride) {
    double2 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) { size_t k = i + (size_t)j * 256; v[j] = k < n ? __ldcs(p + k) : make_double2(0, 0); }
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j].x + v[j].y;
  }
  if (acc == 1.2345) out[0] = acc;
}
__global__ void k_flush_read(const double2* __restrict__ p, size_t n, double* out) {
  double acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += p[i].x;
  if (acc == 1.2345) out[0] = acc;
}
template <int U>
float run(const double2* d, size_t n, double* out, int grid, int mode, const double2* fl, size_t nfl, void* flw) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float tot = 0; int rep = 20;
  for (int w = 0; w < 3; ++w) k_read<U><<<grid, 256>>>(d, n, out);
  for (int r = 0; r < rep; ++r) {
    if (mode == 1) cudaMemsetAsync(flw, r, 512u << 20);
    if (mode == 2) k_flush_read<<<148 * 8, 256>>>(fl, nfl, out);
    cudaEventRecord(e0); k_read<U><<<grid, 256>>>(d, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1); tot += t;
  }
  return tot / rep;
}
int main() {
  size_t bytes = 148u << 20, n = bytes / 16, nfl = (512u << 20) / 16;
  double2 *d, *fl; double* out; void* flw;
  CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&fl, nfl * 16)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&flw, 512u << 20));
  CK(cudaMemset(d, 0, bytes)); CK(cudaMemset(fl, 0, nfl * 16));
  const char* names[3] = {"warm(back-to-back)", "cold(memset flush)", "cold(read flush)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int gm = 1; gm <= 8; gm *= 2) {
      int grid = 148 * gm;
      float t1 = run<1>(d, n, out, grid, mode, fl, nfl, flw), t2 = run<2>(d, n, out, grid, mode, fl, nfl, flw),
            t4 = run<4>(d, n, out, grid, mode, fl, nfl, flw), t8 = run<8>(d, n, out, grid, mode, fl, nfl, flw);
      printf("%-20s grid=148x%d  U=1 %.1f us %.0f GB/s | U=2 %.1f us %.0f GB/s | U=4 %.1f us %.0f GB/s | U=8 %.1f us %.0f GB/s\n", names[mode], gm,
             t1 * 1e3, bytes / t1 / 1e6, t2 * 1e3, bytes / t2 / 1e6, t4 * 1e3, bytes / t4 / 1e6, t8 * 1e3, bytes / t8 / 1e6);
    }
  return 0;
}
