// Micro-benchmark (GPU box only): sustained FP64 FMA rate of the device (is a 5-9 flop/nonzero fp64 SpMV near it?)
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  if (s == 1.2345) out[0] = s;
}
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 1e-3f + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = fmaf(x[c], a, b);
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += x[c];
  if (s == 1.2345f) out[0] = s;
}
int main() {
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * 8.0 * iters;
    printf("DFMA: %.2f ms -> %.2f T FMA/s = %.2f TFLOP/s fp64\n", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    cudaEventRecord(e0); k_ffma<<<blocks, threads>>>((float*)out, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA: %.2f ms -> %.2f T FMA/s = %.2f TFLOP/s fp32\n", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
  }
  return 0;
}
