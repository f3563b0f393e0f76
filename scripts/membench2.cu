// Micro-benchmark (GPU box only), round 2: the one-pass streaming ceiling for an SpMV-sized working set on B200,
// (a) register-staged ld.global.cs (what k_spmv_sell did in round 1) against
// (b) cp.async.bulk (1-D TMA, SASS UBLKCP) into a shared-memory ring with mbarrier completion,
// for 166 MB and 664 MB, cold (L2 flushed by reading 512 MiB) and back to back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/membench2 scripts/membench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e));return 1;}}while(0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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

// ring of STAGES chunks of CHUNK bytes per block; thread 0 issues, all 256 threads consume from shared memory
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(256) k_bulk(const char* __restrict__ p, size_t nbytes, double* out) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t full[STAGES];
  const size_t nchunk = nbytes / CHUNK;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      size_t c = blockIdx.x + (size_t)s * gridDim.x;
      if (c < nchunk) { mbar_expect_tx(&full[s], CHUNK); bulk_g2s(smem + s * CHUNK, p + c * CHUNK, CHUNK, &full[s]); }
    }
  }
  double acc = 0;
  int it = 0;
  for (size_t c = blockIdx.x; c < nchunk; c += gridDim.x, ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const double2* q = reinterpret_cast<const double2*>(smem + s * CHUNK);
#pragma unroll 4
    for (int i = threadIdx.x; i < CHUNK / 16; i += 256) { double2 v = q[i]; acc += v.x + v.y; }
    __syncthreads();
    if (threadIdx.x == 0) {
      size_t cn = c + (size_t)STAGES * gridDim.x;
      if (cn < nchunk) { mbar_expect_tx(&full[s], CHUNK); bulk_g2s(smem + s * CHUNK, p + cn * CHUNK, CHUNK, &full[s]); }
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

__global__ void k_flush_read(const double2* __restrict__ p, size_t n, double* out) {
  double acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += p[i].x;
  if (acc == 1.2345) out[0] = acc;
}

static const double2* g_fl; static size_t g_nfl; static double* g_out;

template <typename F>
float timeit(F launch, int cold) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float tot = 0; const int rep = 20;
  for (int w = 0; w < 3; ++w) launch();
  for (int r = 0; r < rep; ++r) {
    if (cold) k_flush_read<<<148 * 8, 256>>>(g_fl, g_nfl, g_out);
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1); tot += t;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return tot / rep;
}

template <int STAGES, int CHUNK>
void run_bulk(const char* d, size_t bytes, int per_sm, int cold) {
  const int smem = STAGES * CHUNK;
  cudaFuncSetAttribute(k_bulk<STAGES, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148 * per_sm;
  float t = timeit([&] { k_bulk<STAGES, CHUNK><<<grid, 256, smem>>>(d, bytes, g_out); }, cold);
  cudaError_t e = cudaGetLastError();
  printf("  bulk   stages=%d chunk=%5d B blocks/SM=%d (%3d KB in flight/SM): %.1f us %.0f GB/s %s\n", STAGES, CHUNK, per_sm,
         per_sm * smem / 1024, t * 1e3, bytes / t / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int U>
void run_ldg(const double2* d, size_t bytes, int gm, int cold) {
  float t = timeit([&] { k_read<U><<<148 * gm, 256>>>(d, bytes / 16, g_out); }, cold);
  printf("  ld.cs  U=%d grid=148x%d: %.1f us %.0f GB/s\n", U, gm, t * 1e3, bytes / t / 1e6);
}

int main() {
  const size_t nfl = (512u << 20) / 16;
  double2* fl; double* out;
  CK(cudaMalloc(&fl, nfl * 16)); CK(cudaMalloc(&out, 8)); CK(cudaMemset(fl, 0, nfl * 16));
  g_fl = fl; g_nfl = nfl; g_out = out;
  const size_t sizes[2] = {(size_t)166 << 20, (size_t)664 << 20};
  for (int si = 0; si < 2; ++si) {
    const size_t bytes = sizes[si];
    char* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 0, bytes));
    for (int cold = 1; cold >= 0; --cold) {
      printf("%zu MiB, %s\n", bytes >> 20, cold ? "cold (512 MiB read before every launch)" : "back to back");
      run_ldg<4>((const double2*)d, bytes, 4, cold);
      run_ldg<4>((const double2*)d, bytes, 8, cold);
      run_ldg<8>((const double2*)d, bytes, 4, cold);
      run_ldg<8>((const double2*)d, bytes, 8, cold);
      run_bulk<4, 8192>(d, bytes, 1, cold);
      run_bulk<4, 16384>(d, bytes, 1, cold);
      run_bulk<8, 16384>(d, bytes, 1, cold);
      run_bulk<6, 32768>(d, bytes, 1, cold);
      run_bulk<4, 8192>(d, bytes, 2, cold);
      run_bulk<4, 16384>(d, bytes, 2, cold);
      run_bulk<6, 16384>(d, bytes, 2, cold);
      run_bulk<4, 8192>(d, bytes, 4, cold);
      run_bulk<3, 16384>(d, bytes, 4, cold);
    }
    CK(cudaFree(d));
  }
  return 0;
}
