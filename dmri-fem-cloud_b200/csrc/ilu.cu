// ILU(0) preconditioner on the GPU: KrylovSolver("gmres", "ilu") of the comri C++ demo
// (comri/one-comp/fenics-cpp/main.cpp:180-183) and the notebooks' KrylovSolver("bicgstab") with PETSc's default
// preconditioner (ILU(0) in serial).  PETSc's PCILU is third party; its defaults (levels = 0, natural ordering, no
// shift, unit lower factor stored with U in one array) are restated in oracle/bt_oracle.py: ilu0_factor, which pins
// this file (tests/test_gpu_ilu.py: factor values 1e-12, iteration counts, signals 1e-8).
//
// The operator is the COMPLEX matrix A = P + i c J on the scalar CSR pattern -- the reference's real (re,im)-split
// matrix with every 2x2 block [[a,-b],[b,a]] written as a + ib; scalar ILU(0) on the split matrix drops no fill
// inside the (full) blocks, so both factorisations apply the same operator.
//
// Both the factorisation and the two triangular solves are "synchronisation-free": one warp per row, a row spins on
// the ready flags of the rows it depends on (all with a smaller index for L, a larger one for U) and publishes its own
// flag with a release store.  Blocks are dispatched in index order and a row only waits for rows of earlier blocks (or
// earlier warps of its own block), so the wait always ends.  Ready flags carry an epoch, so nothing is reset between
// solves.  The dependency chains make this a latency-bound secondary path (about as many steps as the matrix has
// level sets); the Jacobi path is the fast one.
#include <cstdio>

#include "btfem_internal.cuh"

namespace {

constexpr int ITPB = 256;
constexpr int IWPB = ITPB / 32;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ void wait_flag(const unsigned int* flag, unsigned int epoch) {
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
  } while (v != epoch);
}
__device__ __forceinline__ void set_flag(unsigned int* flag, unsigned int epoch) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
__device__ __forceinline__ double2 warp_sum2(double2 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// A = P + i c J from the un-preconditioned operator pairs (P_k, J_k)
__global__ void k_ilu_load(int64_t nnz, const double2* __restrict__ PJ, double c, double2* __restrict__ lu) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nnz) lu[k] = make_double2(PJ[k].x, c * PJ[k].y);
}

// In-place ILU(0), IKJ order.  Row i = one warp.  For every lower entry (i,c), ascending c: wait for row c,
// l = a_ic / u_cc, then a_ij -= l * u_cj for the entries j > c of row c that exist in row i (lanes over row c's U part).
__global__ void __launch_bounds__(ITPB) k_ilu_factor(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                     const int32_t* __restrict__ diagpos, double2* lu, unsigned int* flag,
                                                     unsigned int epoch) {
  const int i = blockIdx.x * IWPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int r0 = rowptr[i], r1 = rowptr[i + 1], di = diagpos[i];
  for (int k = r0; k < di; ++k) {
    const int c = colidx[k];
    if (lane == 0) wait_flag(flag + c, epoch);
    __syncwarp();
    const int dc = diagpos[c], c1 = rowptr[c + 1];
    volatile double2* vlu = lu;
    double2 a, u;
    a.x = vlu[k].x; a.y = vlu[k].y;
    u.x = vlu[dc].x; u.y = vlu[dc].y;
    const double2 l = cdiv(a, u);
    for (int kk = dc + 1 + lane; kk < c1; kk += 32) {
      const int j = colidx[kk];
      int lo = k + 1, hi = r1;   // columns of row i are sorted: binary search behind position k
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (colidx[mid] < j) lo = mid + 1; else hi = mid;
      }
      if (lo < r1 && colidx[lo] == j) {
        double2 ucj, aij;
        ucj.x = vlu[kk].x; ucj.y = vlu[kk].y;
        aij.x = vlu[lo].x; aij.y = vlu[lo].y;
        const double2 p = cmul(l, ucj);
        vlu[lo].x = aij.x - p.x;
        vlu[lo].y = aij.y - p.y;
      }
    }
    __syncwarp();
    if (lane == 0) { vlu[k].x = l.x; vlu[k].y = l.y; }
    __syncwarp();
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    set_flag(flag + i, epoch);
  }
}

// y = L^-1 v (unit diagonal), rows ascending
__global__ void __launch_bounds__(ITPB) k_ilu_lower(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                    const int32_t* __restrict__ diagpos, const double2* __restrict__ lu,
                                                    const double2* __restrict__ v, double2* y, unsigned int* flag,
                                                    unsigned int epoch) {
  const int i = blockIdx.x * IWPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int r0 = rowptr[i], di = diagpos[i];
  double2 acc = make_double2(0.0, 0.0);
  for (int k = r0 + lane; k < di; k += 32) {
    const int c = colidx[k];
    wait_flag(flag + c, epoch);
    const volatile double2* vy = y;
    double2 yc;
    yc.x = vy[c].x; yc.y = vy[c].y;
    const double2 p = cmul(lu[k], yc);
    acc.x += p.x;
    acc.y += p.y;
  }
  acc = warp_sum2(acc);
  if (lane == 0) {
    const double2 b = v[i];
    y[i] = make_double2(b.x - acc.x, b.y - acc.y);
    __threadfence();
    set_flag(flag + i, epoch);
  }
}

// z = U^-1 y, rows descending (warp w of the grid takes row n - 1 - w)
__global__ void __launch_bounds__(ITPB) k_ilu_upper(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                    const int32_t* __restrict__ diagpos, const double2* __restrict__ lu,
                                                    const double2* __restrict__ y, double2* z, unsigned int* flag,
                                                    unsigned int epoch) {
  const int w = blockIdx.x * IWPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int i = n - 1 - w;
  const int di = diagpos[i], r1 = rowptr[i + 1];
  double2 acc = make_double2(0.0, 0.0);
  for (int k = di + 1 + lane; k < r1; k += 32) {
    const int c = colidx[k];
    wait_flag(flag + c, epoch);
    const volatile double2* vz = z;
    double2 zc;
    zc.x = vz[c].x; zc.y = vz[c].y;
    const double2 p = cmul(lu[k], zc);
    acc.x += p.x;
    acc.y += p.y;
  }
  acc = warp_sum2(acc);
  if (lane == 0) {
    const double2 b = y[i];
    z[i] = cdiv(make_double2(b.x - acc.x, b.y - acc.y), lu[di]);
    __threadfence();
    set_flag(flag + i, epoch);
  }
}

}  // namespace

// (Re)factor A = P + i c J.  d_PJ must hold the un-preconditioned pairs (bt_combine with BTFEM_PC_NONE).
void bt_ilu_factor(btfem* h, double c, cudaStream_t st) {
  BT_REQUIRE(h->nv_own < 0, "ILU(0): whole-mesh handles");
  if (h->ilu_valid && h->ilu_c == c) return;
  const int n = (int)h->ndof;
  h->d_ilu.alloc(h->nnz);
  h->d_ilu_y.alloc(n);
  if (h->d_ilu_flag.n != (size_t)3 * n) {
    h->d_ilu_flag.alloc((size_t)3 * n);
    h->d_ilu_flag.zero(st);
    h->ilu_epoch = 0;
  }
  k_ilu_load<<<(int)((h->nnz + ITPB - 1) / ITPB), ITPB, 0, st>>>(h->nnz, h->d_PJ.p, c, h->d_ilu.p);
  ++h->ilu_epoch;
  k_ilu_factor<<<(n + IWPB - 1) / IWPB, ITPB, 0, st>>>(n, h->d_rowptr.p, h->d_colidx.p, h->d_diagpos.p, h->d_ilu.p,
                                                       h->d_ilu_flag.p, h->ilu_epoch);
  BT_CUDA(cudaGetLastError());
  h->ilu_c = c;
  h->ilu_valid = true;
}

// out = U^-1 L^-1 in   (in and out may be the same array)
void bt_ilu_apply(btfem* h, const double2* in, double2* out, cudaStream_t st) {
  const int n = (int)h->ndof;
  ++h->ilu_epoch;
  const int g = (n + IWPB - 1) / IWPB;
  k_ilu_lower<<<g, ITPB, 0, st>>>(n, h->d_rowptr.p, h->d_colidx.p, h->d_diagpos.p, h->d_ilu.p, in, h->d_ilu_y.p,
                                  h->d_ilu_flag.p + n, h->ilu_epoch);
  k_ilu_upper<<<g, ITPB, 0, st>>>(n, h->d_rowptr.p, h->d_colidx.p, h->d_diagpos.p, h->d_ilu.p, h->d_ilu_y.p, out,
                                  h->d_ilu_flag.p + 2 * (size_t)n, h->ilu_epoch);
  BT_CUDA(cudaGetLastError());
}

// factor values in CSR order (parity hook)
void bt_ilu_get(btfem* h, double* out) {
  BT_REQUIRE(h->ilu_valid, "no ILU(0) factorisation on this handle yet");
  h->d_ilu.download(reinterpret_cast<double2*>(out), h->stream);
}
