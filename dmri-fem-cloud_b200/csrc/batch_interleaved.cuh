// Batches of independent solves on one mesh (HARDI sweeps), part 2: the kernel-chain batch kernels (shared-operator SELL
// kernel, member-interleaved k_hb_*) and the member-interleaved batch as one cooperative launch.
// Included by solve.cu inside its anonymous namespace, behind batch_persistent.cuh.

// ---- batched solves: the shared-operator SELL kernel.  The members of a batch (HARDI: directions x b-values on
// one mesh) differ in the direction g and the scalar c only, so the matrix is read ONCE per slice for BM members:
// (column, P|Q, Jx, Jy, Jz) = 44 bytes per nonzero for the whole group instead of 20 bytes per member, and it
// stays in L2 from one SpMV to the next (31 MB at 46 k vertices, whatever the batch size).  Per member the
// kernel forms J_g,k = (g.J_k)/P_rr with the rounding of k_combine and gathers x_m[col]; the row sums run in
// ascending column order and the dot-product partials follow the static schedule of the single-solve kernel,
// so every member gets the bits of its one-at-a-time solve.  Block = (slices of the schedule, member group).
template <int MODE, int BM>
__global__ void __launch_bounds__(TPB, 2) k_spmv_sell_batch(SpmvArgs a) {
  __shared__ double s_g[BM][4];       // gx, gy, gz, c of the group's members
  __shared__ unsigned int s_mask;     // members that still work
  const int m0 = blockIdx.y * BM;
  const int lane = threadIdx.x & 31;
  const int wpb = TPB / 32;
  if (threadIdx.x < 32) {
    bool act = false;
    if (lane < BM && m0 + lane < a.members) {
      const int m = m0 + lane;
      const KrylovCtrl* ctrl = a.ctrl + m;
      double c;
      if (MODE == MODE_RHS) {
        act = ctrl->failed == 0;
        c = ctrl->theta_cb_scale * a.cb[(size_t)m * a.step_stride + ctrl->step_next];
      } else {
        act = ctrl->done == 0;
        c = ctrl->theta_cA_scale * a.cA[(size_t)m * a.step_stride + ctrl->step];
      }
      s_g[lane][0] = a.gdirs[3 * m];
      s_g[lane][1] = a.gdirs[3 * m + 1];
      s_g[lane][2] = a.gdirs[3 * m + 2];
      s_g[lane][3] = c;
    }
    const unsigned int mask = __ballot_sync(0xffffffffu, act);
    if (lane == 0) s_mask = mask;
  }
  __syncthreads();
  const unsigned int mask = s_mask;
  if (mask == 0) return;
  const double2* __restrict__ x0 =
      (MODE == MODE_RHS || MODE == MODE_RESID) ? a.u : (MODE == MODE_V ? a.p : a.s);
  x0 += (size_t)m0 * a.vec_stride;
  // per-thread dot-product partials of the BM members live in shared memory (registers go to the row sums)
  __shared__ double s_acc[BM][2][TPB];
#pragma unroll
  for (int m = 0; m < BM; ++m) s_acc[m][0][threadIdx.x] = s_acc[m][1][threadIdx.x] = 0.0;
  const int w = blockIdx.x * wpb + (threadIdx.x >> 5);
  const int kend = __ldg(a.sched_ptr + 2 * w + 2);
  for (int k = __ldg(a.sched_ptr + 2 * w); k < kend; ++k) {
    const int slice = __ldg(a.sched + k);
    const int base = __ldg(a.slice_ptr + slice);
    const int width = (__ldg(a.slice_ptr + slice + 1) - base) >> 5;
    const int row = __ldg(a.sell_row + slice * 32 + lane);       // -1: padding slot past the last row
    const double di = row >= 0 ? __ldg(a.dinv + row) : 0.0;
    const int32_t* cp = a.sell_col + base + lane;
    const double2* pq = a.PQs + base + lane;
    const double2* jxy = a.Jxys + base + lane;
    const double* jz = a.Jzs + base + lane;
    double yr[BM], yi[BM];
#pragma unroll
    for (int m = 0; m < BM; ++m) yr[m] = yi[m] = 0.0;
    // software pipeline: the matrix entry of column j+1 is in flight while the BM gathers of column j are
    int col_n = 0;
    double2 pq_n = make_double2(0.0, 0.0), jv_n = pq_n;
    double jz_n = 0.0;
    if (width > 0) { col_n = __ldg(cp); pq_n = __ldg(pq); jv_n = __ldg(jxy); jz_n = __ldg(jz); }
#pragma unroll 1
    for (int j = 0; j < width; ++j) {
      const int col = col_n;
      const double2 pqv = pq_n, jv = jv_n;
      const double jzv = jz_n;
      if (j + 1 < width) {
        col_n = __ldg(cp + (j + 1) * 32);
        pq_n = __ldg(pq + (j + 1) * 32);
        jv_n = __ldg(jxy + (j + 1) * 32);
        jz_n = __ldg(jz + (j + 1) * 32);
      }
      const double pa = MODE == MODE_RHS ? pqv.y : pqv.x;
      const double2* xc = x0 + col;
#pragma unroll
      for (int h0 = 0; h0 < BM; h0 += 4) {     // four gathers in flight at a time
        double2 xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (mask >> (h0 + q) & 1u) xv[q] = ldv_gather_f64x2(xc + (size_t)(h0 + q) * a.vec_stride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = h0 + q;
          if (mask >> m & 1u) {
            const double pb = s_g[m][3] * comb_jg(s_g[m][0], s_g[m][1], s_g[m][2], jv.x, jv.y, jzv, di);
            yr[m] = fma(pa, xv[q].x, yr[m]);
            yr[m] = fma(-pb, xv[q].y, yr[m]);
            yi[m] = fma(pa, xv[q].y, yi[m]);
            yi[m] = fma(pb, xv[q].x, yi[m]);
          }
        }
      }
    }
    if (row >= 0) {
#pragma unroll
      for (int m = 0; m < BM; ++m)
        if (mask >> m & 1u) {
          double t[2] = {s_acc[m][0][threadIdx.x], s_acc[m][1][threadIdx.x]};
          row_epilogue<MODE>(member_at(a, m0 + m), row, make_double2(yr[m], yi[m]), t);
          s_acc[m][0][threadIdx.x] = t[0];
          s_acc[m][1][threadIdx.x] = t[1];
        }
    }
  }
#pragma unroll 1
  for (int m = 0; m < BM; ++m)
    if (mask >> m & 1u) {   // block-uniform: mask is shared
      double t[2] = {s_acc[m][0][threadIdx.x], s_acc[m][1][threadIdx.x]};
      mode_finalize<MODE>(member_at(a, m0 + m), t);
    }
}

// ------------------------------------------------------------------------------------ batched solves, member-interleaved
// HARDI-type sweeps: many (direction, b) solves on ONE mesh.  Members travel in groups of HB = 8 whose Krylov vectors
// are interleaved member-innermost, x[row][m]: the 8 entries of a row are one 128-byte line, so ONE gather serves 8
// members, and the operator is read ONCE per group in its direction-independent form ((P|Q), Jx, Jy, Jz: 44 B per
// nonzero, L2-resident for sweep-sized meshes) instead of 20 B per nonzero per member.  A block works on one SELL
// slice at a time: 32 rows x 8 members = 256 threads, lane = (row & 3, member); the operator loads of a row are
// broadcasts within its 8 lanes.  Every member keeps the arithmetic of its one-at-a-time solve on k_spmv_sell
// (J_g formed with the rounding of k_combine, row sums in ascending column order) and its dot products are reduced in
// a fixed order that does not depend on the batch, so a member gets the same bits in whatever batch it travels.
constexpr int HB = 8;

struct HbLane {
  int m;            // member of this lane inside the group
  int member;       // global member index
  bool act;         // the member exists and still works
  double gx, gy, gz, cc;
};

// per-member sums over the block -> partials[member][q][block]; the last block of the GROUP finishes all 8 members:
// warp w adds member w's partials over the blocks in a fixed order and lane 0 calls fin(member, totals).
template <int NV, typename F>
__device__ __forceinline__ void hb_reduce(const SpmvArgs& a, int m0, double (&v)[NV], int slot, int tk, F&& fin) {
  __shared__ double sm[NV][TPB / 32][HB];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double t = v[q];
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    if (lane < HB) sm[q][warp][lane] = t;
  }
  __syncthreads();
  if (threadIdx.x < HB && m0 + (int)threadIdx.x < a.members) {
    double* part = a.partials + (size_t)(m0 + threadIdx.x) * a.part_stride;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < TPB / 32; ++w) t += sm[q][w][threadIdx.x];
      part[(slot + q) * BT_MAX_PARTIALS + blockIdx.x] = t;
    }
  }
  unsigned int* ticket = &a.ctrl0[m0].ticket[tk];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int member = m0 + warp;   // 8 warps <-> 8 members
  if (member < a.members) {
    const volatile double* part = a.partials + (size_t)member * a.part_stride;
    double tot[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double acc = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) acc += part[(slot + q) * BT_MAX_PARTIALS + b];
      tot[q] = warp_sum(acc);
    }
    if (lane == 0) fin(member, tot);
  }
  __syncthreads();
  if (threadIdx.x == 0) *ticket = 0;
}

// The WHILE condition of the step graph, evaluated by ONE thread in a kernel of its own behind the kernels that
// change ctrl->done (no race between the finishing blocks of different groups; ~2 us per pass).
__global__ void k_hb_cond(SpmvArgs a) {
  if (!a.use_cond) return;
  unsigned int any = 0;
  for (int b = 0; b < a.members; ++b) any |= (a.ctrl0[b].done == 0);
  cudaGraphSetConditional(a.cond, any);
}

__global__ void k_hb_set_ic(int n, size_t group_stride, const double* __restrict__ ic, double2* __restrict__ u) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < (size_t)n * HB) u[blockIdx.y * group_stride + i] = make_double2(ic[i / HB], 0.0);
}

// vectors of group g: u, r, rp, p, v, s, t at k * npad * HB behind a.u + g * 7 * npad * HB (a.vec_stride = 7 * npad)
struct HbVecs {
  double2 *u, *r, *rp, *p, *v, *s, *t;
};
__device__ __forceinline__ HbVecs hb_vecs(const SpmvArgs& a, int g) {
  const size_t npadHB = a.vec_stride / 7 * HB;
  double2* base = a.u + (size_t)g * 7 * npadHB;
  HbVecs w;
  w.u = base; w.r = base + npadHB; w.rp = base + 2 * npadHB; w.p = base + 3 * npadHB; w.v = base + 4 * npadHB;
  w.s = base + 5 * npadHB; w.t = base + 6 * npadHB;
  return w;
}

template <int MODE>
__global__ void __launch_bounds__(TPB, 4) k_hb_spmv(SpmvArgs a) {
  const int g = blockIdx.y, m0 = g * HB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  HbLane L;
  L.m = lane & (HB - 1);
  L.member = m0 + L.m;
  L.act = false;
  L.gx = L.gy = L.gz = L.cc = 0.0;
  if (L.member < a.members) {
    const KrylovCtrl* ctrl = a.ctrl0 + L.member;
    if (MODE == MODE_RHS) {
      L.act = ctrl->failed == 0;
      L.cc = ctrl->theta_cb_scale * a.cb[(size_t)L.member * a.step_stride + ctrl->step_next];
    } else {
      L.act = ctrl->done == 0;
      L.cc = ctrl->theta_cA_scale * a.cA[(size_t)L.member * a.step_stride + ctrl->step];
    }
    L.gx = a.gdirs[3 * L.member]; L.gy = a.gdirs[3 * L.member + 1]; L.gz = a.gdirs[3 * L.member + 2];
  }
  if (__syncthreads_or(L.act) == 0) return;   // the whole group has stopped
  const HbVecs w = hb_vecs(a, g);
  const double2* __restrict__ x = (MODE == MODE_RHS) ? w.u : (MODE == MODE_V ? w.p : w.s);
  const int rslot = warp * 4 + (lane >> 3);   // row slot of this lane inside a slice
  double acc[2] = {0.0, 0.0};
  // a block takes a CONTIGUOUS range of slices: neighbouring rows gather the same x lines (128 bytes per row and
  // group), which then come from this SM's L1 instead of L2
  const int per = (a.nslice + gridDim.x - 1) / gridDim.x;
  const int s_end = min(a.nslice, ((int)blockIdx.x + 1) * per);
  for (int slice = blockIdx.x * per; slice < s_end; ++slice) {
    const int base = __ldg(a.slice_ptr + slice);
    const int width = (__ldg(a.slice_ptr + slice + 1) - base) >> 5;
    const int row = __ldg(a.sell_row + slice * 32 + rslot);   // -1: padding slot past the last row
    const double di = row >= 0 ? __ldg(a.dinv + row) : 0.0;
    const int32_t* cp = a.sell_col + base + rslot;
    const double2* pq = a.PQs + base + rslot;
    const double2* jxy = a.Jxys + base + rslot;
    const double* jz = a.Jzs + base + rslot;
    double yr = 0.0, yi = 0.0;
    constexpr int U = 4;
    for (int j0 = 0; j0 < width; j0 += U) {
      int col[U];
      double pa[U], pb[U];
      double2 xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u;
        col[u] = j < width ? __ldg(cp + j * 32) : -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (col[u] >= 0) xv[u] = ldv_gather_f64x2(x + (size_t)col[u] * HB + L.m);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u;
        if (col[u] >= 0) {
          const double2 pqv = __ldg(pq + j * 32), jv = __ldg(jxy + j * 32);
          const double jzv = __ldg(jz + j * 32);
          pa[u] = MODE == MODE_RHS ? pqv.y : pqv.x;
          pb[u] = L.cc * comb_jg(L.gx, L.gy, L.gz, jv.x, jv.y, jzv, di);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (col[u] >= 0) {
          yr = fma(pa[u], xv[u].x, yr);
          yr = fma(-pb[u], xv[u].y, yr);
          yi = fma(pa[u], xv[u].y, yi);
          yi = fma(pb[u], xv[u].x, yi);
        }
    }
    if (row >= 0 && L.act) {
      const size_t e = (size_t)row * HB + L.m;
      const double2 y = make_double2(yr, yi);
      if (MODE == MODE_RHS) {
        w.r[e] = y; w.rp[e] = y;
        w.p[e] = make_double2(0.0, 0.0);
        w.v[e] = make_double2(0.0, 0.0);
        acc[0] += y.x * y.x + y.y * y.y;
      } else if (MODE == MODE_V) {
        w.v[e] = y;
        const double2 q = w.rp[e];
        acc[0] += y.x * q.x + y.y * q.y;
      } else {
        w.t[e] = y;
        const double2 sv = w.s[e];
        acc[0] += sv.x * y.x + sv.y * y.y;
        acc[1] += y.x * y.x + y.y * y.y;
      }
    }
  }
  if (MODE == MODE_RHS) {
    double v1[1] = {acc[0]};
    hb_reduce<1>(a, m0, v1, 0, TK_RHS, [&](int member, const double (&tot)[1]) {
      KrylovCtrl* ctrl = a.ctrl0 + member;
      if (ctrl->failed) return;
      const double bn = sqrt(tot[0]);
      ctrl->bnorm = bn;
      ctrl->ttol = fmax(ctrl->rtol * bn, ctrl->atol);
      ctrl->rho_old = 1.0; ctrl->alpha = 1.0; ctrl->omega = 1.0;
      ctrl->iters = 0;
      ctrl->step = ctrl->step_next;
      ctrl->step_next = ctrl->step_next + 1;
      ctrl->done = 0; ctrl->reason = 0;
      ctrl->rho = tot[0];
      ctrl->rnorm = bn;
      if (!(bn == bn) || isinf(bn)) { ctrl->done = 1; ctrl->reason = BTFEM_ENAN; }
      else if (bn <= ctrl->ttol) { ctrl->done = 1; ctrl->reason = bn < ctrl->atol ? 3 : 2; }
    });
  } else if (MODE == MODE_V) {
    double v1[1] = {acc[0]};
    hb_reduce<1>(a, m0, v1, 2, TK_V, [&](int member, const double (&tot)[1]) {
      KrylovCtrl* ctrl = a.ctrl0 + member;
      if (ctrl->done) return;
      if (tot[0] == 0.0) { ctrl->done = 1; ctrl->reason = BTFEM_EBREAKDOWN; ctrl->alpha = 0.0; }
      else ctrl->alpha = ctrl->rho / tot[0];
    });
  } else {
    double v2[2] = {acc[0], acc[1]};
    hb_reduce<2>(a, m0, v2, 3, TK_T, [&](int member, const double (&tot)[2]) {
      KrylovCtrl* ctrl = a.ctrl0 + member;
      if (ctrl->done) return;
      ctrl->omega = (tot[1] == 0.0) ? 0.0 : tot[0] / tot[1];
    });
  }
}

// per-thread member constants of the interleaved vector kernels (the grid stride is a multiple of HB)
__device__ __forceinline__ int hb_member(const SpmvArgs& a) { return blockIdx.y * HB + (threadIdx.x & (HB - 1)); }

// p <- r - omega*beta*v + beta*p
__global__ void __launch_bounds__(TPB) k_hb_update_p(SpmvArgs a) {
  const int member = hb_member(a);
  bool act = false;
  double beta = 0.0, ob = 0.0;
  if (member < a.members) {
    const KrylovCtrl* ctrl = a.ctrl0 + member;
    act = ctrl->done == 0;
    if (act) {
      beta = (ctrl->rho / ctrl->rho_old) * (ctrl->alpha / ctrl->omega);
      ob = ctrl->omega * beta;
    }
  }
  if (__syncthreads_or(act) == 0) return;
  const HbVecs w = hb_vecs(a, blockIdx.y);
  const size_t ne = (size_t)a.n * HB;
  if (act)
    for (size_t i = blockIdx.x * (size_t)TPB + threadIdx.x; i < ne; i += (size_t)gridDim.x * TPB) {
      const double2 rr = w.r[i], vv = w.v[i];
      double2 pp = w.p[i];
      pp.x = rr.x - ob * vv.x + beta * pp.x;
      pp.y = rr.y - ob * vv.y + beta * pp.y;
      w.p[i] = pp;
    }
}

// s <- r - alpha*v
__global__ void __launch_bounds__(TPB) k_hb_update_s(SpmvArgs a) {
  const int member = hb_member(a);
  bool act = false;
  double alpha = 0.0;
  if (member < a.members) {
    const KrylovCtrl* ctrl = a.ctrl0 + member;
    act = ctrl->done == 0;
    alpha = ctrl->alpha;
  }
  if (__syncthreads_or(act) == 0) return;
  const HbVecs w = hb_vecs(a, blockIdx.y);
  const size_t ne = (size_t)a.n * HB;
  if (act)
    for (size_t i = blockIdx.x * (size_t)TPB + threadIdx.x; i < ne; i += (size_t)gridDim.x * TPB) {
      const double2 rr = w.r[i], vv = w.v[i];
      w.s[i] = make_double2(rr.x - alpha * vv.x, rr.y - alpha * vv.y);
    }
}

// x <- x + alpha*p + omega*s ; r <- s - omega*t ; rho' = (r,rp) ; ||r|| ; convergence test   (per member)
__global__ void __launch_bounds__(TPB) k_hb_update_xr(SpmvArgs a) {
  const int member = hb_member(a);
  bool act = false, fresh = false;
  double alpha = 0.0, omega = 0.0;
  if (member < a.members) {
    const KrylovCtrl* ctrl = a.ctrl0 + member;
    act = ctrl->done == 0;
    alpha = ctrl->alpha;
    omega = ctrl->omega;
    fresh = ctrl->iters == 0;   // zero initial guess: x starts from 0
  }
  if (__syncthreads_or(act) == 0) return;
  const HbVecs w = hb_vecs(a, blockIdx.y);
  const size_t ne = (size_t)a.n * HB;
  double acc[2] = {0.0, 0.0};
  if (act)
    for (size_t i = blockIdx.x * (size_t)TPB + threadIdx.x; i < ne; i += (size_t)gridDim.x * TPB) {
      const double2 pp = w.p[i], ss = w.s[i], tt = w.t[i], q = w.rp[i];
      double2 xx = fresh ? make_double2(0.0, 0.0) : w.u[i];
      xx.x += alpha * pp.x + omega * ss.x;
      xx.y += alpha * pp.y + omega * ss.y;
      w.u[i] = xx;
      const double2 rr = make_double2(ss.x - omega * tt.x, ss.y - omega * tt.y);
      w.r[i] = rr;
      acc[0] += rr.x * q.x + rr.y * q.y;
      acc[1] += rr.x * rr.x + rr.y * rr.y;
    }
  hb_reduce<2>(a, blockIdx.y * HB, acc, 5, TK_XR, [&](int mb, const double (&tot)[2]) {
    KrylovCtrl* ctrl = a.ctrl0 + mb;
    if (ctrl->done) return;
    const double rho_used = ctrl->rho, om = ctrl->omega;
    ctrl->rho_old = rho_used;
    ctrl->rho = tot[0];
    const double dp = sqrt(tot[1]);
    ctrl->rnorm = dp;
    const int it = ctrl->iters + 1;
    ctrl->iters = it;
    if (!(dp == dp) || isinf(dp)) { ctrl->done = 1; ctrl->reason = BTFEM_ENAN; }
    else if (dp <= ctrl->ttol) { ctrl->done = 1; ctrl->reason = dp < ctrl->atol ? 3 : 2; }
    else if (dp >= ctrl->dtol * ctrl->bnorm) { ctrl->done = 1; ctrl->reason = BTFEM_EDTOL; }
    else if (rho_used == 0.0 || om == 0.0) { ctrl->done = 1; ctrl->reason = BTFEM_EBREAKDOWN; }
    else if (it >= ctrl->maxit) { ctrl->done = 1; ctrl->reason = BTFEM_ENOTCONV; }
  });
}

// end of a time step: statistics; "converged before the first iteration with a zero guess returns x = 0"
__global__ void __launch_bounds__(TPB) k_hb_step_end(SpmvArgs a) {
  const int member = hb_member(a);
  bool zero = false;
  if (member < a.members) {
    KrylovCtrl* ctrl = a.ctrl0 + member;
    if (!ctrl->failed) {
      const int it = ctrl->iters;
      zero = it == 0 && ctrl->reason > 0;
      if (blockIdx.x == 0 && threadIdx.x < HB) {
        ctrl->total_iters += it;
        if (it > ctrl->max_iters) ctrl->max_iters = it;
      }
    }
  }
  if (zero) {
    const HbVecs w = hb_vecs(a, blockIdx.y);
    const size_t ne = (size_t)a.n * HB;
    for (size_t i = blockIdx.x * (size_t)TPB + threadIdx.x; i < ne; i += (size_t)gridDim.x * TPB)
      w.u[i] = make_double2(0.0, 0.0);
  }
}

// signal = sum_i lumped_i * Re u_i, split by compartment, per member
__global__ void __launch_bounds__(TPB) k_hb_signal(SpmvArgs a, const double* __restrict__ lumped,
                                                   const int32_t* __restrict__ comp) {
  const HbVecs w = hb_vecs(a, blockIdx.y);
  const size_t ne = (size_t)a.n * HB;
  double acc[2] = {0.0, 0.0};
  for (size_t i = blockIdx.x * (size_t)TPB + threadIdx.x; i < ne; i += (size_t)gridDim.x * TPB) {
    const size_t row = i / HB;
    const double v = lumped[row] * w.u[i].x;
    if (comp[row] == 0) acc[0] += v; else acc[1] += v;
  }
  hb_reduce<2>(a, blockIdx.y * HB, acc, 6, TK_SIG, [&](int mb, const double (&tot)[2]) {
    a.sig_out[2 * mb] = tot[0];
    a.sig_out[2 * mb + 1] = tot[1];
  });
}

// ---- the member-interleaved batch as ONE cooperative kernel (BTFEM_BATCH_PERSIST=hb).  The passes of the member
// layout are bound by the L1 wavefront rate: a 16-byte gather per lane touches ~32 different 128-byte lines per
// warp-load (measured 1.6 clk per nonzero and member, k_bicgstab_coop_batch).  Here a warp works on 4 rows x 8 members:
// x[row][member] is one full line per row, so a warp-load touches 4 lines; the direction-independent operator
// ((P|Q), Jx, Jy, Jz) is read once per 8 members and J_g formed per member with the rounding of k_combine (k_hb_spmv).
// Same phases, grid barrier and reductions as k_bicgstab_coop_batch; work units of a pass are "super-tasks" of 4
// neighbouring slices (32 warps x 4 rows) dealt round-robin over the blocks, so that every block sees long and short
// rows of the sorting windows alike.
__device__ __forceinline__ void chb_rebuild(PbState& S, int M) {
  int na = 0, ng = 0;
  for (int m = 0; m < M; ++m)
    if ((S.active >> m) & 1u) S.act[na++] = (unsigned char)m;
  for (int g = 0; g * HB < M; ++g)
    if ((S.active >> (g * HB)) & 0xffu) S.actg[ng++] = (unsigned char)g;
  S.nact = na;
  S.ngact = ng;
}

// acc[gi][q]: this thread's terms for member (lane & 7) of active group gi
template <int NT>
__device__ __forceinline__ void chb_grid_reduce(const SpmvArgs& a, PbState& S, double (*wacc)[2][2][HB], const double (&acc)[2][2],
                                                int M, int nq, int slot, GridSync& g) {
  constexpr int NWB = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int gi = 0; gi < 2; ++gi)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      double t = acc[gi][q];
      t += __shfl_xor_sync(0xffffffffu, t, 8);
      t += __shfl_xor_sync(0xffffffffu, t, 16);
      if (lane < HB) wacc[warp][gi][q][lane] = t;
    }
  __syncthreads();
  if ((int)threadIdx.x < 2 * HB * nq) {   // block partial of (active group gi, member ml of it, q)
    const int gi = threadIdx.x / (HB * nq), r = threadIdx.x - gi * HB * nq, ml = r / nq, q = r - ml * nq;
    if (gi < S.ngact) {
      const int m = S.actg[gi] * HB + ml;
      if (m < M && ((S.active >> m) & 1u)) {
        double t = 0.0;
        for (int w = 0; w < NWB; ++w) t += wacc[w][gi][q][ml];
        __stcg(a.partials + (size_t)m * a.part_stride + (size_t)(slot + q) * BT_MAX_PARTIALS + blockIdx.x, t);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) g.arrive_wait();
  __syncthreads();
  const int nval = S.nact * nq;
  for (int i = warp; i < nval; i += NWB) {
    const int m = S.act[i / nq], q = i % nq;
    const double* pp = a.partials + (size_t)m * a.part_stride + (size_t)(slot + q) * BT_MAX_PARTIALS;
    double t = 0.0;
    for (unsigned int b = lane; b < gridDim.x; b += 32) t += __ldcg(pp + b);
    t = warp_sum(t);
    if (lane == 0) S.tot[2 * m + q] = t;
  }
  __syncthreads();
}

template <int NT, int U>
__global__ void __launch_bounds__(NT, 1) k_bicgstab_coop_hb(SpmvArgs a) {
  __shared__ PbState S;
  __shared__ double wacc[NT / 32][2][2][HB];
  static_assert(NT % 256 == 0, "a super-task is NT / 256 slices: 8 warps x 4 rows each");
  constexpr int SPT = NT / 256;   // slices per super-task
  const int M = a.pb.members;
  for (int m = 0; m < M; ++m)
    if (a.ctrl0[m].failed) return;   // uniform: written by an earlier launch only
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ml = lane & (HB - 1);
  const unsigned int all = M >= 32 ? 0xffffffffu : (1u << M) - 1u;
  const int n = a.n, NS = a.nslice, NB = gridDim.x;
  const size_t npadHB = a.vec_stride / 7 * HB;
  KrylovCtrl* ctrl0 = a.ctrl0;
  auto load_step_scalars = [&](int step) {   // thread 0
    for (int m = 0; m < M; ++m) {
      S.cb[m] = ctrl0->theta_cb_scale * a.cb[(size_t)m * a.step_stride + step];
      S.cA[m] = ctrl0->theta_cA_scale * a.cA[(size_t)m * a.step_stride + step];
    }
  };
  if (threadIdx.x == 0) {
    S.mode = MODE_RHSP;
    S.step = a.step_begin;
    S.it = 0;
    S.fail = 0;
    S.active = all;
    for (int m = 0; m < M; ++m) {
      S.total_iters[m] = ctrl0[m].total_iters;
      S.max_iters[m] = ctrl0[m].max_iters;
      S.reason[m] = 0;
      S.its[m] = 0;
      S.rho[m] = S.rho_old[m] = S.alpha[m] = S.omega[m] = 1.0;
      S.beta[m] = 0.0;
      S.bn[m] = S.ttol[m] = S.rnorm[m] = 0.0;
      S.gd[m][0] = a.gdirs[3 * m]; S.gd[m][1] = a.gdirs[3 * m + 1]; S.gd[m][2] = a.gdirs[3 * m + 2];
    }
    S.bar_target = a.gridbar[32];
    if (a.step_begin < a.step_end) load_step_scalars(a.step_begin);
    chb_rebuild(S, M);
  }
  __syncthreads();
  GridSync gs;
  gs.count = a.gridbar;
  gs.target = 0;
  const bool prof_on = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long tprev = prof_on ? global_ns() : 0ull;
#define PROF(k)                                   \
  if (prof_on) {                                  \
    const unsigned long long tn_ = global_ns();   \
    a.prof[k] += tn_ - tprev;                     \
    tprev = tn_;                                  \
  }
  const double atol = ctrl0->atol, rtol = ctrl0->rtol, dtol = ctrl0->dtol;
  const int maxit = ctrl0->maxit;

  for (;;) {
    const int mode = S.mode, step = S.step;
    if (step >= a.step_end || S.fail) break;
    const int nact = S.nact, ngact = S.ngact;
    const unsigned int active = S.active;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    // rows of the (active group, row) space this block takes in the vector phases (8 members = one line per row)
    const long long WV = (long long)ngact * n;
    const long long rlo = WV * blockIdx.x / NB, rhi = WV * (blockIdx.x + 1) / NB;
    const unsigned long long tp0 = a.prof ? global_ns() : 0ull;
    // ---- pass over the super-tasks (active group, 4 slices) of this block
    {
      const int NST = (NS + SPT - 1) / SPT;
      const int rslot = (warp & 7) * 4 + (lane >> 3);
      for (int j = blockIdx.x; j < ngact * NST; j += NB) {
        const int gi = j / NST, st = j - gi * NST, g = S.actg[gi];
        const int slice = SPT * st + (warp >> 3);
        if (slice >= NS) continue;
        const int m = g * HB + ml;
        const bool act = m < M && ((active >> m) & 1u);
        const int mm = m < M ? m : 0;
        const double gx = S.gd[mm][0], gy = S.gd[mm][1], gz = S.gd[mm][2];
        const double cc = mode == MODE_RHSP ? S.cb[mm] : S.cA[mm];
        double2* vb = a.u + (size_t)g * 7 * npadHB;   // u, r, rp, p, v, s, t of the group, npadHB apart
        const double2* x = (mode == MODE_RHSP ? vb : (mode == MODE_V ? vb + 3 * npadHB : vb + 5 * npadHB)) + ml;
        const int base = __ldg(a.slice_ptr + slice);
        const int width = (__ldg(a.slice_ptr + slice + 1) - base) >> 5;
        const int row = __ldg(a.sell_row + slice * 32 + rslot);   // -1: padding slot past the last row
        const double di = row >= 0 ? __ldg(a.dinv + row) : 0.0;
        const size_t e = (size_t)(row >= 0 ? row : 0) * HB + ml;
        double2 op = make_double2(0.0, 0.0);
        if (row >= 0 && act && mode != MODE_RHSP) op = (mode == MODE_V ? vb + 2 * npadHB : vb + 5 * npadHB)[e];
        const int32_t* cp = a.sell_col + base + rslot;
        const double2* pq = a.PQs + base + rslot;
        const double2* jxy = a.Jxys + base + rslot;
        const double* jz = a.Jzs + base + rslot;
        double yr = 0.0, yi = 0.0;
        // Every load of a round is issued before any is used (volatile asm: ptxas otherwise sinks each operator load
        // next to its use -- ncu on the first version: one L2 round trip per COLUMN behind the gather), and the
        // columns of the next round travel with them.
        int col[U];
#pragma unroll
        for (int u = 0; u < U; ++u) col[u] = u < width ? ldv_nc_i32(cp + u * 32) : -1;
        for (int j0 = 0; j0 < width; j0 += U) {
          double2 xv[U], pqv[U], jv[U];
          double jzv[U];
          int coln[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (col[u] >= 0) xv[u] = ldv_gather_f64x2(x + (size_t)col[u] * HB);
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (col[u] >= 0) {
              pqv[u] = ldv_nc_f64x2(pq + (j0 + u) * 32);
              jv[u] = ldv_nc_f64x2(jxy + (j0 + u) * 32);
              jzv[u] = ldv_nc_f64(jz + (j0 + u) * 32);
            }
#pragma unroll
          for (int u = 0; u < U; ++u) coln[u] = j0 + U + u < width ? ldv_nc_i32(cp + (j0 + U + u) * 32) : -1;
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (col[u] >= 0) {
              const double pa = mode == MODE_RHSP ? pqv[u].y : pqv[u].x;
              const double pb = cc * comb_jg(gx, gy, gz, jv[u].x, jv[u].y, jzv[u], di);
              yr = fma(pa, xv[u].x, yr);
              yr = fma(-pb, xv[u].y, yr);
              yi = fma(pa, xv[u].y, yi);
              yi = fma(pb, xv[u].x, yi);
            }
#pragma unroll
          for (int u = 0; u < U; ++u) col[u] = coln[u];
        }
        if (row >= 0 && act) {
          const double2 y = make_double2(yr, yi);
          double t0, t1 = 0.0;
          if (mode == MODE_V) {
            vb[4 * npadHB + e] = y;
            t0 = y.x * op.x + y.y * op.y;
          } else if (mode == MODE_T) {
            vb[6 * npadHB + e] = y;
            t0 = op.x * y.x + op.y * y.y;
            t1 = y.x * y.x + y.y * y.y;
          } else {
            vb[npadHB + e] = y;
            vb[2 * npadHB + e] = y;
            t0 = y.x * y.x + y.y * y.y;
          }
          if (gi == 0) { acc[0][0] += t0; acc[0][1] += t1; }
          else { acc[1][0] += t0; acc[1][1] += t1; }
        }
      }
    }
    if (a.prof && threadIdx.x == 0) a.prof[16 + blockIdx.x] += global_ns() - tp0;   // per-block time inside the passes
    PROF(mode == MODE_RHSP ? 0 : (mode == MODE_V ? 3 : 7));
    gs.target = S.bar_target;
    if (mode == MODE_V) {
      // ---- alpha = rho / (r^, v) ; s = r - alpha v
      chb_grid_reduce<NT>(a, S, wacc, acc, M, 1, 2, gs);
      PROF(4);
      if ((int)threadIdx.x < nact) {
        const int m = S.act[threadIdx.x];
        const double d = S.tot[2 * m];
        if (d == 0.0) { S.reason[m] = BTFEM_EBREAKDOWN; S.its[m] = S.it; atomicMin(&S.fail, (int)BTFEM_EBREAKDOWN); }
        else S.alpha[m] = S.rho[m] / d;
      }
      __syncthreads();
      if (!S.fail) {
        for (long long i = rlo * HB + threadIdx.x; i < rhi * HB; i += NT) {
          const long long rr_ = i / HB;
          const int gi = (int)(rr_ / n), g = S.actg[gi], m = g * HB + ml;
          if (m < M && ((active >> m) & 1u)) {
            const size_t e = (size_t)(i - (long long)gi * n * HB);
            double2* vb = a.u + (size_t)g * 7 * npadHB;
            const double alpha = S.alpha[m];
            const double2 rr = vb[npadHB + e], vv = vb[4 * npadHB + e];
            vb[5 * npadHB + e] = make_double2(rr.x - alpha * vv.x, rr.y - alpha * vv.y);
          }
        }
        PROF(5);
        grid_barrier(gs);
        PROF(6);
        if (threadIdx.x == 0) { S.mode = MODE_T; S.bar_target = gs.target; }
        __syncthreads();
        continue;
      }
    } else if (mode == MODE_T) {
      // ---- omega = (t,s) / (t,t) ; x <- x + alpha p + omega s ; r <- s - omega t ; rho' = (r, r^) ; ||r||
      chb_grid_reduce<NT>(a, S, wacc, acc, M, 2, 3, gs);
      PROF(8);
      if ((int)threadIdx.x < nact) {
        const int m = S.act[threadIdx.x];
        S.omega[m] = (S.tot[2 * m + 1] == 0.0) ? 0.0 : S.tot[2 * m] / S.tot[2 * m + 1];
      }
      __syncthreads();
      acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
      {
        const bool fresh = S.it == 0;   // zero initial guess: x starts from 0
        for (long long i = rlo * HB + threadIdx.x; i < rhi * HB; i += NT) {
          const long long rr_ = i / HB;
          const int gi = (int)(rr_ / n), g = S.actg[gi], m = g * HB + ml;
          if (m < M && ((active >> m) & 1u)) {
            const size_t e = (size_t)(i - (long long)gi * n * HB);
            double2* vb = a.u + (size_t)g * 7 * npadHB;
            const double alpha = S.alpha[m], omega = S.omega[m];
            const double2 pp = vb[3 * npadHB + e], ss = vb[5 * npadHB + e], tt = vb[6 * npadHB + e], qq = vb[2 * npadHB + e];
            double2 xx = make_double2(0.0, 0.0);
            if (!fresh) xx = vb[e];
            xx.x += alpha * pp.x + omega * ss.x;
            xx.y += alpha * pp.y + omega * ss.y;
            vb[e] = xx;
            const double2 rr = make_double2(ss.x - omega * tt.x, ss.y - omega * tt.y);
            vb[npadHB + e] = rr;
            const double t0 = rr.x * qq.x + rr.y * qq.y, t1 = rr.x * rr.x + rr.y * rr.y;
            if (gi == 0) { acc[0][0] += t0; acc[0][1] += t1; }
            else { acc[1][0] += t0; acc[1][1] += t1; }
          }
        }
      }
      PROF(9);
      chb_grid_reduce<NT>(a, S, wacc, acc, M, 2, 5, gs);
      PROF(10);
      if (threadIdx.x < 32) {   // warp 0, one lane per active member: the convergence tests of KSPConvergedDefault
        const int it = S.it + 1;
        const unsigned int before = S.active;
        __syncwarp();
        if (lane < nact) {
          const int m = S.act[lane];
          const double rho_used = S.rho[m], omega = S.omega[m], rho_new = S.tot[2 * m], rnorm = sqrt(S.tot[2 * m + 1]);
          S.rho_old[m] = rho_used;
          S.rho[m] = rho_new;
          S.rnorm[m] = rnorm;
          int reason = 0;
          if (!(rnorm == rnorm) || isinf(rnorm)) reason = BTFEM_ENAN;
          else if (rnorm <= S.ttol[m]) reason = rnorm < atol ? 3 : 2;
          else if (rnorm >= dtol * S.bn[m]) reason = BTFEM_EDTOL;
          else if (rho_used == 0.0 || omega == 0.0) reason = BTFEM_EBREAKDOWN;
          else if (it >= maxit) reason = BTFEM_ENOTCONV;
          if (reason != 0) {
            S.reason[m] = reason;
            S.its[m] = it;
            atomicAnd(&S.active, ~(1u << m));
            if (reason < 0) atomicMin(&S.fail, reason);
          } else {
            S.beta[m] = (rho_new / rho_used) * (S.alpha[m] / omega);
          }
        }
        __syncwarp();
        if (lane == 0) {
          S.it = it;
          if (S.active != before) chb_rebuild(S, M);
        }
      }
      __syncthreads();
    } else {
      // ---- ||r|| of every member: start of the Krylov solves of this step
      chb_grid_reduce<NT>(a, S, wacc, acc, M, 1, 0, gs);
      PROF(10);
      if (threadIdx.x < 32) {
        const unsigned int before = S.active;
        __syncwarp();
        if (lane < nact) {
          const int m = S.act[lane];
          const double bn = sqrt(S.tot[2 * m]);
          const double ttol = fmax(rtol * bn, atol);
          S.bn[m] = bn;
          S.ttol[m] = ttol;
          S.rho[m] = S.tot[2 * m]; S.rho_old[m] = 1.0; S.alpha[m] = 1.0; S.omega[m] = 1.0; S.rnorm[m] = bn;
          int reason = 0;
          if (!(bn == bn) || isinf(bn)) reason = BTFEM_ENAN;
          else if (bn <= ttol) reason = bn < atol ? 3 : 2;
          if (reason != 0) {
            S.reason[m] = reason;
            S.its[m] = 0;
            atomicAnd(&S.active, ~(1u << m));
            if (reason < 0) atomicMin(&S.fail, reason);
          }
        }
        __syncwarp();
        if (lane == 0) {
          S.it = 0;
          if (S.active != before) chb_rebuild(S, M);
        }
      }
      __syncthreads();
    }
    if (!S.fail && S.active != 0u) {
      // ---- p <- r - omega*beta*v + beta*p   (first iteration: p = r), then v = A p; chunks cut from the groups that go on
      const unsigned int active2 = S.active;
      const long long WV2 = (long long)S.ngact * n;
      const long long plo = WV2 * blockIdx.x / NB, phi = WV2 * (blockIdx.x + 1) / NB;
      const bool first = S.it == 0;
      for (long long i = plo * HB + threadIdx.x; i < phi * HB; i += NT) {
        const long long rr_ = i / HB;
        const int gi = (int)(rr_ / n), g = S.actg[gi], m = g * HB + ml;
        if (m < M && ((active2 >> m) & 1u)) {
          const size_t e = (size_t)(i - (long long)gi * n * HB);
          double2* vb = a.u + (size_t)g * 7 * npadHB;
          const double2 rr = vb[npadHB + e];
          if (first) {
            vb[3 * npadHB + e] = rr;
          } else {
            const double beta = S.beta[m], ob = S.omega[m] * beta;
            const double2 vv = vb[4 * npadHB + e];
            double2 pp = vb[3 * npadHB + e];
            pp.x = rr.x - ob * vv.x + beta * pp.x;
            pp.y = rr.y - ob * vv.y + beta * pp.y;
            vb[3 * npadHB + e] = pp;
          }
        }
      }
      PROF(1);
      grid_barrier(gs);
      PROF(2);
      if (threadIdx.x == 0) { S.mode = MODE_V; S.bar_target = gs.target; }
      __syncthreads();
      continue;
    }
    // ---- every member has finished the time step (or one has failed, which ends the batch)
    if (!S.fail) {
      bool anyz = false;   // converged before the first iteration with a zero guess: PETSc returns x = 0
      for (int m = 0; m < M; ++m)
        if (S.its[m] == 0 && S.reason[m] > 0) {
          anyz = true;
          double2* u = a.u + (size_t)(m / HB) * 7 * npadHB + (m % HB);
          for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += NB * NT) u[(size_t)i * HB] = make_double2(0.0, 0.0);
        }
      if (anyz) grid_barrier(gs);   // the next right-hand sides gather x
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int fail = S.fail;
      for (int m = 0; m < M; ++m) {
        const int it = S.its[m];
        S.total_iters[m] += it;
        S.max_iters[m] = max(S.max_iters[m], it);
        if (blockIdx.x == 0) {
          KrylovCtrl* ctrl = ctrl0 + m;
          ctrl->bnorm = S.bn[m]; ctrl->ttol = S.ttol[m]; ctrl->rnorm = S.rnorm[m];
          ctrl->rho = S.rho[m]; ctrl->rho_old = S.rho_old[m]; ctrl->alpha = S.alpha[m]; ctrl->omega = S.omega[m];
          ctrl->iters = it; ctrl->reason = S.reason[m]; ctrl->done = 1;
          ctrl->step = step; ctrl->step_next = step + 1;
          ctrl->total_iters = S.total_iters[m]; ctrl->max_iters = S.max_iters[m];
          if (fail) ctrl->failed = fail;   // a failure of any member stops every member (k_step_fail)
        }
      }
      S.bar_target = gs.target;
      if (!fail) {
        S.step = step + 1;
        S.mode = MODE_RHSP;
        S.it = 0;
        S.active = all;
        for (int m = 0; m < M; ++m) { S.reason[m] = 0; S.its[m] = 0; }
        if (step + 1 < a.step_end) load_step_scalars(step + 1);
        chb_rebuild(S, M);
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.gridbar[32] = S.bar_target;
#undef PROF
}
