// Peer-memory signalling latency between two B200s (NVLink 5 / NVSwitch), the primitives libbtfem's
// row-partitioned solve is built from.  One process, two devices, peer access enabled.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/p2p_latency scripts/p2p_latency.cu
// Prints the round-trip time of a flag ping-pong for several load/store flavours, and the time of one
// "payload + fence + flag" exchange (the in-kernel all-reduce pattern).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long* p) {
  return *(volatile const unsigned long long*)p; }
__device__ __forceinline__ void st_rel(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_rlx(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_vol(unsigned long long* p, unsigned long long v) {
  *(volatile unsigned long long*)p = v; }

// mode 0: acquire/release.sys   1: relaxed.sys ld/st   2: volatile ld/st   3: volatile + __threadfence_system before the store
template <int MODE>
__global__ void pingpong(int me, unsigned long long* mine, unsigned long long* peer, int n, unsigned long long base) {
  for (int i = 1; i <= n; ++i) {
    const unsigned long long s = base + i;
    if (me == 0) {
      if (MODE == 0) st_rel(peer, s); else if (MODE == 1) st_rlx(peer, s); else { if (MODE == 3) __threadfence_system(); st_vol(peer, s); }
      while ((MODE == 0 ? ld_acq(mine) : MODE == 1 ? ld_rlx(mine) : ld_vol(mine)) < s) {}
    } else {
      while ((MODE == 0 ? ld_acq(mine) : MODE == 1 ? ld_rlx(mine) : ld_vol(mine)) < s) {}
      if (MODE == 0) st_rel(peer, s); else if (MODE == 1) st_rlx(peer, s); else { if (MODE == 3) __threadfence_system(); st_vol(peer, s); }
    }
  }
}

// the all-reduce pattern: both sides write a payload to the peer, fence, publish, wait for the peer's flag, read
// VARIANT 0: __threadfence_system + st.release   1: payload and flag in ONE 16-byte store (value, seq), no fence
__global__ void exchange(int me, unsigned long long* mine, unsigned long long* peer, int n, unsigned long long base,
                         double* out, int variant) {
  double acc = 0;
  for (int i = 1; i <= n; ++i) {
    const unsigned long long s = base + i;
    const int buf = (int)(s & 1);
    if (variant == 0) {
      *(volatile double*)(peer + 8 + buf) = (double)i;
      __threadfence_system();
      st_rel(peer + buf, s);
      while (ld_acq(mine + buf) < s) {}
      acc += *(volatile double*)(mine + 8 + buf);
    } else {
      // 16-byte vector store: (seq, value) lands atomically in practice (one 16 B aligned write)
      double v = (double)i;
      asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(peer + 16 + 2 * buf), "l"(s), "l"(__double_as_longlong(v)) : "memory");
      unsigned long long fs, fv;
      do {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(fs), "=l"(fv) : "l"(mine + 16 + 2 * buf) : "memory");
      } while (fs < s);
      acc += __longlong_as_double(fv);
    }
  }
  out[0] = acc;
}

// bulk halo push: nthreads x 16 B to the peer, every thread fences, last thread publishes; the peer waits
__global__ void push(int me, unsigned long long* mine, unsigned long long* peer, double2* peer_buf, int count, int n,
                     unsigned long long base, int fence_each) {
  __shared__ int dummy;
  for (int i = 1; i <= n; ++i) {
    const unsigned long long s = base + i;
    for (int e = threadIdx.x; e < count; e += blockDim.x) peer_buf[e] = make_double2((double)i, (double)e);
    if (fence_each) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      st_rel(peer, s);
      while (ld_acq(mine) < s) {}
    }
    __syncthreads();
  }
  (void)dummy;
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  unsigned long long* flag[2];
  double* out[2];
  double2* buf[2];
  cudaStream_t st[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&flag[d], 4096));
    CK(cudaMemset(flag[d], 0, 4096));
    CK(cudaMalloc(&out[d], 64));
    CK(cudaMalloc(&buf[d], 1 << 22));
    CK(cudaStreamCreate(&st[d]));
  }
  for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
  const int n = 20000;
  unsigned long long base = 0;
  cudaEvent_t e0, e1;
  CK(cudaSetDevice(0));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, auto launch, double per, int iters) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaSetDevice(0));
      CK(cudaEventRecord(e0, st[0]));
      launch(0);
      CK(cudaEventRecord(e1, st[0]));
      CK(cudaSetDevice(1));
      launch(1);
      CK(cudaSetDevice(0));
      CK(cudaEventSynchronize(e1));
      CK(cudaSetDevice(1));
      CK(cudaStreamSynchronize(st[1]));
      base += n;
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1) printf("%-58s %8.3f us per %s\n", name, 1e3 * ms / iters, per > 1 ? "exchange" : "round trip");
    }
  };
#define PP(MODE, NAME) run(NAME, [&](int d) { pingpong<MODE><<<1, 1, 0, st[d]>>>(d, flag[d], flag[1 - d], n, base); }, 1, n)
  PP(0, "ping-pong  st.release.sys / ld.acquire.sys");
  PP(1, "ping-pong  st.relaxed.sys / ld.relaxed.sys");
  PP(2, "ping-pong  volatile st / volatile ld");
  PP(3, "ping-pong  __threadfence_system + volatile st / volatile ld");
  run("exchange   payload, __threadfence_system, release flag", [&](int d) { exchange<<<1, 1, 0, st[d]>>>(d, flag[d], flag[1 - d], n, base, out[d], 0); }, 2, n);
  run("exchange   (seq,value) in one 16-byte relaxed store", [&](int d) { exchange<<<1, 1, 0, st[d]>>>(d, flag[d], flag[1 - d], n, base, out[d], 1); }, 2, n);
  for (int count : {256, 4096, 16384}) {
    char nm[128];
    snprintf(nm, sizeof nm, "push %5d x 16 B, fence per thread, flag, wait (1 block)", count);
    run(nm, [&](int d) { push<<<1, 1024, 0, st[d]>>>(d, flag[d], flag[1 - d], buf[1 - d], count, n / 10, base, 1); }, 2, n / 10);
  }
  return 0;
}
