// Batches of independent solves on one mesh (HARDI sweeps), part 1: the whole time loop of a lock-step batch as ONE
// cooperative launch -- the TMA-ring kernel and the many-warp kernel on the SELL copies (the default).
// Included by solve.cu inside its anonymous namespace, behind the single-solve persistent kernel (uses SpmvArgs, WarpRing,
// GridSync, grid_barrier, warp_sum, slice_product and the MODE_* enum defined there).

// ---- lock-step batches (HARDI sweeps: directions x b-values on one mesh) as ONE persistent kernel.
// Same machinery as k_bicgstab_persistent -- per-warp TMA rings, grid barrier, every block finishes the reductions --
// with every phase covering all members that still iterate, so that the barrier / reduction latency (15 us per
// iteration, as much as the whole iteration of ONE member on a 46 k-vertex mesh) is paid once per batch iteration:
//  * passes: members with the same gradient direction share one operator stream (A = P + i c J_g: b only enters
//    through c).  A warp fetches a piece once and runs it for every member of the group ("units": the gathers of the
//    next unit are in flight while one unit is multiplied), so the stream is read once per group and the per-piece
//    costs (mbarrier wait, list entry, column decode) are spread over the group;
//  * dot products: every thread keeps its per-member partial sums in shared memory, the block / grid reduction visits
//    them in the order of the single-solve kernel -> a member gets the bits of its one-at-a-time persistent solve;
//  * vector phases: one sweep over (member, row) with batched loads; thread -> row mapping of the single-solve kernel.
// Members iterate in lock-step inside a time step (one common iteration counter); a member that has converged
// drops out of the passes and phases until the next step starts.
constexpr int PB_NW = 8, PB_D = 3;
constexpr int pb_smem() { return ps_smem(PB_NW, PB_D) + 2 * PB_MAX * PB_NW * 32 * 8; }

struct PbState {
  double rho[CB_MAX], rho_old[CB_MAX], alpha[CB_MAX], omega[CB_MAX], beta[CB_MAX];
  double bn[CB_MAX], ttol[CB_MAX], rnorm[CB_MAX], cA[CB_MAX], cb[CB_MAX];
  double tot[2 * CB_MAX];
  double wsum[2 * CB_MAX][PB_NW];
  long long total_iters[CB_MAX];
  int max_iters[CB_MAX], its[CB_MAX], reason[CB_MAX];
  unsigned int active;                  // bit m: member m still iterates in this time step
  unsigned char gmask[CB_MAX];          // per group: its active members (bit j: member g_m0 + j)
  unsigned char unit[CB_MAX][PB_GM];    // per group: the member behind unit j (its j-th active member)
  unsigned char act[CB_MAX];            // active members, ascending
  int nact;
  int cut[2], cut_nact;                 // coop kernel: first and past-the-last item of this block; the nact they belong to
  unsigned char actg[CB_MAX];           // member-interleaved form: groups of 8 members with an active member
  int ngact;
  double gd[CB_MAX][3];                 // member-interleaved form: gradient direction of every member
  int it, mode, step, fail;
  unsigned int bar_target;
};

// thread 0: the lists that follow from S.active
__device__ __forceinline__ void pb_rebuild(PbState& S, const PbArgs& pb) {
  int na = 0;
  for (int g = 0; g < pb.groups; ++g) {
    const int m0 = pb.g_m0[g], nm = pb.g_nm[g];
    const unsigned int gm = (S.active >> m0) & ((1u << nm) - 1u);
    S.gmask[g] = (unsigned char)gm;
    int j = 0;
    for (int k = 0; k < nm; ++k)
      if ((gm >> k) & 1u) {
        S.unit[g][j++] = (unsigned char)(m0 + k);
        S.act[na++] = (unsigned char)(m0 + k);
      }
  }
  S.nact = na;
}
__device__ __forceinline__ int pb_next_group(const PbState& S, const PbArgs& pb, int g) {
  for (int k = g + 1; k < pb.groups; ++k)
    if (S.gmask[k]) return k;
  return -1;
}

// gather step of one unit: columns of ring piece c (landed), x of member `m`
template <int D>
__device__ __forceinline__ void pb_gather(const WarpRing<D>& r, unsigned int c, const int4 d, const double2* __restrict__ x,
                                          const double2* __restrict__ opv, double2 (&xv)[BT_PS_W], double2& op) {
  const unsigned char* sp = r.stage(c);
  const int lane = threadIdx.x & 31;
  if (r.colu == 9) {
    const uint16_t* cs = reinterpret_cast<const uint16_t*>(sp) + lane;
    const int ref = d.w >> 1;
#pragma unroll
    for (int j = 0; j < BT_PS_W; ++j)
      if (j < d.y) xv[j] = ldv_gather_f64x2(x + (ref + (int)cs[j * 32]));
  } else {
    const int32_t* cs = reinterpret_cast<const int32_t*>(sp) + lane;
#pragma unroll
    for (int j = 0; j < BT_PS_W; ++j)
      if (j < d.y) xv[j] = ldv_gather_f64x2(x + cs[j * 32]);
  }
  if (d.w & 1) {
    const int row = reinterpret_cast<const int32_t*>(sp + (size_t)d.y * r.colu * 64)[lane];
    op = make_double2(0.0, 0.0);
    if (row >= 0 && opv) op = opv[row];
  }
}
template <int D>
__device__ __forceinline__ void pb_fma(const WarpRing<D>& r, unsigned int c, const int4 d, double cc,
                                       const double2 (&xv)[BT_PS_W], double& ar, double& ai) {
  const double2* vs = reinterpret_cast<const double2*>(r.stage(c) + d.y * (r.colu == 9 ? 64 : 128)) + (threadIdx.x & 31);
#pragma unroll
  for (int j = 0; j < BT_PS_W; ++j)
    if (j < d.y) {
      const double2 val = vs[j * 32];
      const double pa = val.x, pb = cc * val.y;
      ar = fma(pa, xv[j].x, ar);
      ar = fma(-pb, xv[j].y, ar);
      ai = fma(pa, xv[j].y, ai);
      ai = fma(pb, xv[j].x, ai);
    }
}

// One pass of this warp over its list for the `na` units (active members) of a group: ring pieces c0 .. c0 + np - 1 of
// stream T, refilled like stream_pass (the first entries of `Tnext` follow those of T).  acc: this thread's column of
// the per-member accumulators, acc[(2 m + q) * NT].
template <int D, int NT>
__device__ __forceinline__ unsigned int pb_stream_pass(const SpmvArgs& a, int mode, const WarpRing<D>& r, unsigned int c0,
                                                       const unsigned char* T, const unsigned char* Tnext,
                                                       const unsigned char* unit, int na, const double* ccs, double* acc) {
  const int lane = threadIdx.x & 31;
  const int np = r.np;
  const int nnext = min(D, np);
  if (lane == 0)
    for (int i = 0; i < np && np + i < D; ++i) r.fetch(Tnext, i, c0 + np + i);
  if (np == 0) return c0;
  const double2* xbase = mode == MODE_RHSP ? a.u : (mode == MODE_V ? a.p : a.s);
  const double2* opbase = mode == MODE_V ? a.rp : (mode == MODE_T ? a.s : nullptr);
  const size_t vs = a.vec_stride;
  auto entry = [&](int i) { return __ldg(r.pieces + (i < np ? i : i - np)); };
  double ar[PB_GM], ai[PB_GM];
#pragma unroll
  for (int j = 0; j < PB_GM; ++j) ar[j] = ai[j] = 0.0;
  double2 xv[2][BT_PS_W], op[2];
  op[0] = op[1] = make_double2(0.0, 0.0);
  int4 d = __ldg(r.pieces), dn = d, dr = make_int4(0, 0, 0, 0);
  if (np > 1) dn = __ldg(r.pieces + 1);
  if (D < np + nnext) dr = entry(D);
  const bool even = (na & 1) == 0;
  auto gather_unit = [&](unsigned int c, const int4 dd, int j, double2 (&xb)[BT_PS_W], double2& ob) {
    const size_t off = (size_t)unit[j] * vs;
    pb_gather(r, c, dd, xbase + off, opbase ? opbase + off : nullptr, xb, ob);
  };
  r.wait(c0);
  gather_unit(c0, d, 0, xv[0], op[0]);
  for (int k = 0; k < np; ++k) {
    const unsigned int c = c0 + k;
    const bool last = (d.w & 1) != 0;
    int row = -1;
    if (last) row = reinterpret_cast<const int32_t*>(r.stage(c) + (size_t)d.y * r.colu * 64)[lane];
#pragma unroll
    for (int j = 0; j < PB_GM; ++j)
      if (j < na) {
        // the next unit's loads go out before this unit's arithmetic: the next member on this piece, or (even unit
        // counts: the register buffers then alternate across the piece boundary) the first member on the next piece
        if (j + 1 < na) {
          gather_unit(c, d, j + 1, xv[(j + 1) & 1], op[(j + 1) & 1]);
        } else if (even && k + 1 < np) {
          r.wait(c + 1);
          gather_unit(c + 1, dn, 0, xv[0], op[0]);
        }
        const int m = unit[j];
        pb_fma(r, c, d, ccs[m], xv[j & 1], ar[j], ai[j]);
        if (last) {
          if (row >= 0) {
            const size_t off = (size_t)m * vs + row;
            const double2 y = make_double2(ar[j], ai[j]), o = op[j & 1];
            double* a0 = acc + (size_t)(2 * m) * NT;
            if (mode == MODE_V) {
              a.v[off] = y;
              a0[0] += y.x * o.x + y.y * o.y;
            } else if (mode == MODE_T) {
              a.t[off] = y;
              a0[0] += o.x * y.x + o.y * y.y;
              a0[NT] += y.x * y.x + y.y * y.y;
            } else {
              a.r[off] = y;
              a.rp[off] = y;
              a0[0] += y.x * y.x + y.y * y.y;
            }
          }
          ar[j] = 0.0;
          ai[j] = 0.0;
        }
      }
    // piece k is done for every unit: its stage takes the entry D pieces further (see stream_pass)
    __syncwarp();
    const int nx = k + D;
    if (lane == 0 && nx < np + nnext) {
      if (a.ps_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      r.fetch_desc(nx < np ? T : Tnext, dr, c0 + nx);
    }
    if (nx + 1 < np + nnext) dr = entry(nx + 1);
    if (!even && k + 1 < np) {
      r.wait(c + 1);
      gather_unit(c + 1, dn, 0, xv[0], op[0]);
    }
    d = dn;
    if (k + 2 < np) dn = __ldg(r.pieces + k + 2);
  }
  return c0 + np;
}

// Sums of the per-thread accumulators acc_s[(2 m + q) * NT + thread], q < nq, of the active members over the whole
// grid -> S.tot[2 m + q] in every block.  Order of the additions: that of grid_reduce.
template <int NW>
__device__ __forceinline__ void pb_grid_reduce(const SpmvArgs& a, PbState& S, const double* acc_s, int nq, int slot,
                                               GridSync& g) {
  constexpr int NT = NW * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nact = S.nact, nval = nact * nq;
  for (int i = 0; i < nval; ++i) {
    const int v = 2 * S.act[i / nq] + (i % nq);
    const double t = warp_sum(acc_s[(size_t)v * NT + threadIdx.x]);
    if (lane == 0) S.wsum[v][warp] = t;
  }
  __syncthreads();
  for (int i = warp; i < nval; i += NW) {
    const int m = S.act[i / nq], q = i % nq;
    double t = lane < NW ? S.wsum[2 * m + q][lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) __stcg(a.partials + (size_t)m * a.part_stride + (size_t)(slot + q) * BT_MAX_PARTIALS + blockIdx.x, t);
  }
  __syncthreads();
  if (threadIdx.x == 0) g.arrive_wait();
  __syncthreads();
  for (int i = warp; i < nval; i += NW) {
    const int m = S.act[i / nq], q = i % nq;
    const double* pp = a.partials + (size_t)m * a.part_stride + (size_t)(slot + q) * BT_MAX_PARTIALS;
    double t = 0.0;
    for (unsigned int b = lane; b < gridDim.x; b += 32) t += __ldcg(pp + b);
    t = warp_sum(t);
    if (lane == 0) S.tot[2 * m + q] = t;
  }
  __syncthreads();
}

// Vector phases over (active member, row): work item w of a thread = (member act[w / R], row gid + (w % R) * gsz),
// R = rows per thread -- the thread -> row mapping of pv_update_*; PB_U items are loaded before any is used.
constexpr int PB_U = 4, PB_UV = 8;
__device__ __noinline__ void pb_update_p(int n, size_t vstride, const PbState& S, int gid, int gsz, bool first,
                                         const double2* rv, const double2* v, double2* p) {
  const int R = (n + gsz - 1) / gsz, items = S.nact * R;
  for (int w0 = 0; w0 < items; w0 += PB_UV) {
    double2 rr[PB_UV], vv[PB_UV], pp[PB_UV];
    size_t idx[PB_UV];
    int mm[PB_UV];
#pragma unroll
    for (int u = 0; u < PB_UV; ++u) {
      const int w = w0 + u;
      mm[u] = -1;
      rr[u] = vv[u] = pp[u] = make_double2(0.0, 0.0);
      if (w < items) {
        const int m = S.act[w / R], row = gid + (w % R) * gsz;
        if (row < n) {
          mm[u] = m;
          idx[u] = (size_t)m * vstride + row;
          rr[u] = rv[idx[u]];
          if (!first) { vv[u] = v[idx[u]]; pp[u] = p[idx[u]]; }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PB_UV; ++u)
      if (mm[u] >= 0) {
        if (first) {
          p[idx[u]] = rr[u];
        } else {
          const double beta = S.beta[mm[u]], ob = S.omega[mm[u]] * beta;
          pp[u].x = rr[u].x - ob * vv[u].x + beta * pp[u].x;
          pp[u].y = rr[u].y - ob * vv[u].y + beta * pp[u].y;
          p[idx[u]] = pp[u];
        }
      }
  }
}
__device__ __noinline__ void pb_update_s(int n, size_t vstride, const PbState& S, int gid, int gsz, const double2* rv,
                                         const double2* v, double2* sv) {
  const int R = (n + gsz - 1) / gsz, items = S.nact * R;
  for (int w0 = 0; w0 < items; w0 += PB_UV) {
    double2 rr[PB_UV], vv[PB_UV];
    size_t idx[PB_UV];
    int mm[PB_UV];
#pragma unroll
    for (int u = 0; u < PB_UV; ++u) {
      const int w = w0 + u;
      mm[u] = -1;
      rr[u] = vv[u] = make_double2(0.0, 0.0);
      if (w < items) {
        const int m = S.act[w / R], row = gid + (w % R) * gsz;
        if (row < n) {
          mm[u] = m;
          idx[u] = (size_t)m * vstride + row;
          rr[u] = rv[idx[u]];
          vv[u] = v[idx[u]];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PB_UV; ++u)
      if (mm[u] >= 0) {
        const double alpha = S.alpha[mm[u]];
        sv[idx[u]] = make_double2(rr[u].x - alpha * vv[u].x, rr[u].y - alpha * vv[u].y);
      }
  }
}
// x <- x + alpha p + omega s ; r <- s - omega t ; acc[2m] += (r, r^), acc[2m+1] += (r, r)   (acc zeroed by the caller)
template <int NT>
__device__ __noinline__ void pb_update_xr(int n, size_t vstride, const PbState& S, int gid, int gsz, bool fresh,
                                          const double2* p, const double2* sv, const double2* t, const double2* rp,
                                          double2* x, double2* rv, double* acc) {
  const int R = (n + gsz - 1) / gsz, items = S.nact * R;
  for (int w0 = 0; w0 < items; w0 += PB_U) {
    double2 pp[PB_U], ss[PB_U], tt[PB_U], qq[PB_U], xx[PB_U];
    size_t idx[PB_U];
    int mm[PB_U];
#pragma unroll
    for (int u = 0; u < PB_U; ++u) {
      const int w = w0 + u;
      mm[u] = -1;
      pp[u] = ss[u] = tt[u] = qq[u] = xx[u] = make_double2(0.0, 0.0);
      if (w < items) {
        const int m = S.act[w / R], row = gid + (w % R) * gsz;
        if (row < n) {
          mm[u] = m;
          idx[u] = (size_t)m * vstride + row;
          pp[u] = p[idx[u]]; ss[u] = sv[idx[u]]; tt[u] = t[idx[u]]; qq[u] = rp[idx[u]];
          if (!fresh) xx[u] = x[idx[u]];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PB_U; ++u)
      if (mm[u] >= 0) {
        const double alpha = S.alpha[mm[u]], omega = S.omega[mm[u]];
        xx[u].x += alpha * pp[u].x + omega * ss[u].x;
        xx[u].y += alpha * pp[u].y + omega * ss[u].y;
        x[idx[u]] = xx[u];
        const double2 rr = make_double2(ss[u].x - omega * tt[u].x, ss[u].y - omega * tt[u].y);
        rv[idx[u]] = rr;
        double* a0 = acc + (size_t)(2 * mm[u]) * NT;
        a0[0] += rr.x * qq[u].x + rr.y * qq[u].y;
        a0[NT] += rr.x * rr.x + rr.y * rr.y;
      }
  }
}

template <int NW, int D>
__global__ void __launch_bounds__(NW * 32, 1) k_bicgstab_persistent_batch(SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char ps_ring[];
  __shared__ PbState S;
  constexpr int NT = NW * 32;
  constexpr int ring_bytes = NW * D * PS_STAGE + NW * D * 8;   // ps_smem(NW, D)
  double* acc_s = reinterpret_cast<double*>(ps_ring + ring_bytes);
  double* acc = acc_s + threadIdx.x;
  const int M = a.pb.members;
  for (int m = 0; m < M; ++m)
    if (a.ctrl[m].failed) return;   // uniform: written by an earlier launch only
  const int lane = threadIdx.x & 31;
  const int gsz = gridDim.x * NT, gid = blockIdx.x * NT + threadIdx.x;
  const unsigned int all = M >= 32 ? 0xffffffffu : (1u << M) - 1u;
  auto load_step_scalars = [&](int step) {   // thread 0
    for (int m = 0; m < M; ++m) {
      S.cb[m] = a.ctrl->theta_cb_scale * a.cb[(size_t)m * a.step_stride + step];
      S.cA[m] = a.ctrl->theta_cA_scale * a.cA[(size_t)m * a.step_stride + step];
    }
  };
  if (threadIdx.x == 0) {
    S.mode = MODE_RHSP;
    S.step = a.step_begin;
    S.it = 0;
    S.fail = 0;
    S.active = all;
    for (int m = 0; m < M; ++m) {
      S.total_iters[m] = a.ctrl[m].total_iters;
      S.max_iters[m] = a.ctrl[m].max_iters;
      S.reason[m] = 0;
      S.its[m] = 0;
      S.rho[m] = S.rho_old[m] = S.alpha[m] = S.omega[m] = 1.0;
      S.beta[m] = 0.0;
      S.bn[m] = S.ttol[m] = S.rnorm[m] = 0.0;
    }
    S.bar_target = a.gridbar[32];
    if (a.step_begin < a.step_end) load_step_scalars(a.step_begin);
    pb_rebuild(S, a.pb);
  }
  WarpRing<D> r;
  r.setup(ps_ring, a.ps_ptr, a.ps_piece, a.ps_c16);
  const int nfl = min(D, r.np);
  unsigned int c = 0;
  const size_t sstr = a.pb.stream_stride;
  const unsigned char* inring = a.QJt + (size_t)a.pb.g_dir[0] * sstr;   // the stream whose first nfl pieces are in the ring
  if (lane == 0)
    for (int i = 0; i < nfl; ++i) r.fetch(inring, i, (unsigned int)i);
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

  for (;;) {
    const int mode = S.mode, step = S.step;
    if (step >= a.step_end || S.fail) break;
    const int nact = S.nact;
    for (int i = 0; i < nact; ++i) {
      const int m = S.act[i];
      acc[(size_t)(2 * m) * NT] = 0.0;
      acc[(size_t)(2 * m + 1) * NT] = 0.0;
    }
    // ---- passes: one per group with active members
    {
      const unsigned char* base = mode == MODE_RHSP ? a.QJt : a.PJt;
      const double* ccs = mode == MODE_RHSP ? S.cb : S.cA;
      const int gfirst = pb_next_group(S, a.pb, -1);
      int g = gfirst;
      while (g >= 0) {
        const int gn = pb_next_group(S, a.pb, g);
        const unsigned char* T = base + (size_t)a.pb.g_dir[g] * sstr;
        // after the last group: the first group's v = A p / t = A s pass follows unless the step ends or members drop out
        const unsigned char* Tn = gn >= 0 ? base + (size_t)a.pb.g_dir[gn] * sstr : a.PJt + (size_t)a.pb.g_dir[gfirst] * sstr;
        if (inring != T) {   // the ring holds pieces of another stream: let them land, fetch the right ones
          for (int i = 0; i < nfl; ++i) r.wait(c + i);
          __syncwarp();
          c += nfl;
          if (lane == 0)
            for (int i = 0; i < nfl; ++i) r.fetch(T, i, c + i);
        }
        c = pb_stream_pass<D, NT>(a, mode, r, c, T, Tn, S.unit[g], __popc((unsigned int)S.gmask[g]), ccs, acc);
        inring = Tn;
        g = gn;
      }
    }
    PROF(mode == MODE_RHSP ? 0 : (mode == MODE_V ? 3 : 7));
    gs.target = S.bar_target;
    if (mode == MODE_V) {
      // ---- alpha = rho / (r^, v) ; s = r - alpha v
      pb_grid_reduce<NW>(a, S, acc_s, 1, 2, gs);
      PROF(4);
      if (threadIdx.x < nact) {   // one lane per active member (nact <= PB_MAX <= 32)
        const int m = S.act[threadIdx.x];
        const double d = S.tot[2 * m];
        if (d == 0.0) { S.reason[m] = BTFEM_EBREAKDOWN; S.its[m] = S.it; atomicMin(&S.fail, (int)BTFEM_EBREAKDOWN); }
        else S.alpha[m] = S.rho[m] / d;
      }
      __syncthreads();
      if (!S.fail) {
        pb_update_s(a.n, a.vec_stride, S, gid, gsz, a.r, a.v, a.s);
        PROF(5);
        grid_barrier(gs);
        PROF(6);
        if (threadIdx.x == 0) { S.mode = MODE_T; S.bar_target = gs.target; }
        __syncthreads();
        continue;
      }
    } else if (mode == MODE_T) {
      // ---- omega = (t,s) / (t,t) ; x <- x + alpha p + omega s ; r <- s - omega t ; rho' = (r, r^) ; ||r||
      pb_grid_reduce<NW>(a, S, acc_s, 2, 3, gs);
      PROF(8);
      if (threadIdx.x < nact) {
        const int m = S.act[threadIdx.x];
        S.omega[m] = (S.tot[2 * m + 1] == 0.0) ? 0.0 : S.tot[2 * m] / S.tot[2 * m + 1];
      }
      for (int i = 0; i < nact; ++i) {
        const int m = S.act[i];
        acc[(size_t)(2 * m) * NT] = 0.0;
        acc[(size_t)(2 * m + 1) * NT] = 0.0;
      }
      __syncthreads();
      pb_update_xr<NT>(a.n, a.vec_stride, S, gid, gsz, S.it == 0, a.p, a.s, a.t, a.rp, a.u, a.r, acc);
      PROF(9);
      pb_grid_reduce<NW>(a, S, acc_s, 2, 5, gs);
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
          else if (rnorm <= S.ttol[m]) reason = rnorm < a.ctrl->atol ? 3 : 2;
          else if (rnorm >= a.ctrl->dtol * S.bn[m]) reason = BTFEM_EDTOL;
          else if (rho_used == 0.0 || omega == 0.0) reason = BTFEM_EBREAKDOWN;
          else if (it >= a.ctrl->maxit) reason = BTFEM_ENOTCONV;
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
          if (S.active != before) pb_rebuild(S, a.pb);
        }
      }
      __syncthreads();
    } else {
      // ---- ||r|| of every member: start of the Krylov solves of this step
      pb_grid_reduce<NW>(a, S, acc_s, 1, 0, gs);
      PROF(10);
      if (threadIdx.x < 32) {
        const unsigned int before = S.active;
        __syncwarp();
        if (lane < nact) {
          const int m = S.act[lane];
          const double atol = a.ctrl->atol;
          const double bn = sqrt(S.tot[2 * m]);
          const double ttol = fmax(a.ctrl->rtol * bn, atol);
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
          if (S.active != before) pb_rebuild(S, a.pb);
        }
      }
      __syncthreads();
    }
    if (!S.fail && S.active != 0u) {
      // ---- p <- r - omega*beta*v + beta*p   (first iteration: p = r), then v = A p
      pb_update_p(a.n, a.vec_stride, S, gid, gsz, S.it == 0, a.r, a.v, a.p);
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
          double2* u = a.u + (size_t)m * a.vec_stride;
          for (int i = gid; i < a.n; i += gsz) u[i] = make_double2(0.0, 0.0);
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
          KrylovCtrl* ctrl = a.ctrl + m;
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
        pb_rebuild(S, a.pb);
      }
    }
    __syncthreads();
  }
  // nothing may still be in flight into shared memory when the block retires
  for (int i = 0; i < nfl; ++i) r.wait(c + i);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.gridbar[32] = S.bar_target;
#undef PROF
}


// ---- lock-step batches on a SMALL mesh as one cooperative kernel with many warps ("coop" batch kernel, the default for
// batches of up to CB_MAX members on whole-mesh handles).  ncu on the kernel chain (profiles/r2ai_*) shows where a
// 16-member iteration on a 46 k-vertex mesh goes: 52 us per batched SpMV although neither the operator bytes (one copy
// per direction changes nothing), nor the gather locality (the SELL window changes nothing) bound it -- every block
// walks ctrl -> schedule -> slice -> columns -> gather -> epilogue -> ticket for ONE slice per warp, 6.5 waves of
// blocks per launch.  Here the whole time loop is one launch of 1024-thread blocks (32 warps per SM, <= 64 registers):
//  * passes: the (active member, slice) items of a phase, member-major, are cut into one contiguous chunk per block
//    (<= 2 members per block); inside a block the 32 warps take 32 neighbouring slices at a time (they share gathered
//    x lines in L1), the position rotating from round to round so that no warp always gets the longest rows of a
//    sorting window.  One operator copy per direction (bt_combine: members of a direction share it), plain SELL loads;
//  * vector phases: the (active member, row) space cut into one chunk per block, fully coalesced;
//  * reductions: per-block partials for its <= 2 members, then every block adds the partials of all blocks in a fixed
//    order (deterministic for a given batch; the grouping depends on which members are active, so a member's bits
//    depend on the batch it travels in -- unlike the kernel chain).
// The scalar recurrences, convergence tests and reason codes are those of k_bicgstab_persistent.
constexpr int CB_NT = 1024;
constexpr int CB_CH = 4;                       // columns of a slice per staged piece
constexpr int CB_STAGE = CB_CH * (128 + 512);  // 32 lanes x (4-byte column + 16-byte value pair) per column
constexpr int CB_C0 = 6;                       // cost of an item besides its columns (in columns), for the chunk cuts
constexpr int cb_smem(int NT) { return NT / 32 * 2 * CB_STAGE; }
constexpr int CHB_U = 2;   // columns of a row in flight in k_bicgstab_coop_hb (default variant: 1024 threads, 64 registers)

// thread 0: the list of active members
__device__ __forceinline__ void cb_rebuild(PbState& S, int M) {
  int na = 0;
  for (int m = 0; m < M; ++m)
    if ((S.active >> m) & 1u) S.act[na++] = (unsigned char)m;
  S.nact = na;
}

// acc0 / acc1: this thread's terms for the block's first member (active index ai0) and the one after it; nq values each.
// wm != null: the terms are per-warp sums wm[warp][member][q] already (the passes); the thread terms are not used
template <int NT>
__device__ __forceinline__ void cb_grid_reduce(const SpmvArgs& a, PbState& S, double (*wacc)[4], const double (&acc0)[2],
                                               const double (&acc1)[2], int ai0, int nq, int slot, GridSync& g,
                                               const double (*wm)[CB_MAX][2] = nullptr) {
  constexpr int NWB = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nact = S.nact, nval = nact * nq;
  if (!wm) {
    const double t0 = warp_sum(acc0[0]), t2 = warp_sum(acc1[0]);
    double t1 = 0.0, t3 = 0.0;
    if (nq == 2) { t1 = warp_sum(acc0[1]); t3 = warp_sum(acc1[1]); }
    if (lane == 0) { wacc[warp][0] = t0; wacc[warp][1] = t1; wacc[warp][2] = t2; wacc[warp][3] = t3; }
  }
  __syncthreads();
  if ((int)threadIdx.x < nval) {   // block partial of value (ai, q): zero unless the block holds rows of that member
    const int ai = threadIdx.x / nq, q = threadIdx.x % nq, m = S.act[ai];
    double t = 0.0;
    if (wm) {
      for (int w = 0; w < NWB; ++w) t += wm[w][m][q];
    } else if (ai == ai0 || ai == ai0 + 1) {
      const int c = (ai - ai0) * 2 + q;
      for (int w = 0; w < NWB; ++w) t += wacc[w][c];
    }
    __stcg(a.partials + (size_t)m * a.part_stride + (size_t)(slot + q) * BT_MAX_PARTIALS + blockIdx.x, t);
  }
  __syncthreads();
  if (threadIdx.x == 0) g.arrive_wait();
  __syncthreads();
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

template <int NT>
__global__ void __launch_bounds__(NT, 1) k_bicgstab_coop_batch(SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char cb_stage[];   // [warps][2 stages][CB_STAGE]
  __shared__ PbState S;
  __shared__ double wacc[NT / 32][4];
  __shared__ double wm[NT / 32][CB_MAX][2];   // passes: per-warp sums of the dot-product terms of every member
  __shared__ int sdir[CB_MAX];
  unsigned long long l2_first;                // the operator streams past once per pass: evict-first in L2
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_first));
  constexpr int NWB = NT / 32;
  const int M = a.pb.members;
  for (int m = 0; m < M; ++m)
    if (a.ctrl[m].failed) return;   // uniform: written by an earlier launch only
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int all = M >= 32 ? 0xffffffffu : (1u << M) - 1u;
  const int n = a.n, NS = a.nslice, NB = gridDim.x;
  const size_t vs = a.vec_stride;
  auto load_step_scalars = [&](int step) {   // thread 0
    for (int m = 0; m < M; ++m) {
      S.cb[m] = a.ctrl->theta_cb_scale * a.cb[(size_t)m * a.step_stride + step];
      S.cA[m] = a.ctrl->theta_cA_scale * a.cA[(size_t)m * a.step_stride + step];
    }
  };
  if (threadIdx.x == 0) {
    S.mode = MODE_RHSP;
    S.step = a.step_begin;
    S.it = 0;
    S.fail = 0;
    S.active = all;
    for (int m = 0; m < M; ++m) {
      S.total_iters[m] = a.ctrl[m].total_iters;
      S.max_iters[m] = a.ctrl[m].max_iters;
      S.reason[m] = 0;
      S.its[m] = 0;
      S.rho[m] = S.rho_old[m] = S.alpha[m] = S.omega[m] = 1.0;
      S.beta[m] = 0.0;
      S.bn[m] = S.ttol[m] = S.rnorm[m] = 0.0;
    }
    S.bar_target = a.gridbar[32];
    S.cut_nact = -1;
    if (a.step_begin < a.step_end) load_step_scalars(a.step_begin);
    cb_rebuild(S, M);
  }
  if ((int)threadIdx.x < M) sdir[threadIdx.x] = a.member_dir ? a.member_dir[threadIdx.x] : (int)threadIdx.x;
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

  for (;;) {
    const int mode = S.mode, step = S.step;
    if (step >= a.step_end || S.fail) break;
    const int nact = S.nact;
    // chunk of this block in the (active member, row) space of the vector phases
    const long long WV = (long long)nact * n;
    const long long clo = WV * blockIdx.x / NB, chi = WV * (blockIdx.x + 1) / NB;
    const int aiv0 = (int)(clo / n);
    double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
    int ai0;
    const unsigned long long tp0 = a.prof ? global_ns() : 0ull;
    // ---- pass: y_m = (V.x + i c_m V.y) x_m over this block's chunk of the (active member, slice) items.
    // Columns and values of the NEXT piece (CB_CH columns of a slice; the next slice when this one ends) are on their
    // way into this warp's shared-memory stage (cp.async, every lane its own entries) while the gathers of the current
    // piece are in flight: per piece a warp waits for one memory round trip (the gather) instead of a chain of three
    // (slice -> columns -> gather); ncu on the plain-load version: 36 % of the issue slots wait on those loads.
    {
      // Items are SLICE-major: a block owns a contiguous range of slices (cut once per launch, equal cost) and takes
      // every active member through it, 32 neighbouring (slice, member) items at a time.  The members of a direction
      // read the same operator slice at about the same time (one L2 -> L1 transfer serves them), and each gathered
      // vector is live only in the row band of the block's slices -- with member-major chunks a 16-member pass had all
      // operator copies and all gathered vectors in use at once (134 MB: ncu showed the gathers waiting on DRAM).
      if (S.cut_nact < 0) {   // (uniform)
        __syncthreads();
        if (threadIdx.x < 2) {
          // cost prefix over the slices: a.cb_cost (L1 wavefronts of a slice, counted on the host from its columns), or
          // in closed form columns + CB_C0 per slice
          const long long* pc = a.cb_cost;
          const long long ptot = pc ? __ldg(pc + NS) : (long long)(__ldg(a.slice_ptr + NS) >> 5) + (long long)CB_C0 * NS;
          const long long t = ptot * (blockIdx.x + threadIdx.x) / NB;
          int lo = 0, hi = NS;   // smallest s with prefix(s) >= t
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const long long pm = pc ? __ldg(pc + mid) : (long long)(__ldg(a.slice_ptr + mid) >> 5) + (long long)CB_C0 * mid;
            if (pm >= t) hi = mid; else lo = mid + 1;
          }
          S.cut[threadIdx.x] = lo;
        }
        __syncthreads();
        if (threadIdx.x == 0) S.cut_nact = 0;
      }
      const int slo = S.cut[0], nitems = (S.cut[1] - slo) * nact;
      for (int i = lane; i < 2 * CB_MAX; i += 32) (&wm[warp][0][0])[i] = 0.0;   // this warp's per-member sums
      __syncwarp();
      ai0 = 0;
      const double2* Vb = mode == MODE_RHSP ? a.QJs : a.PJs;
      const double2* xbase = mode == MODE_RHSP ? a.u : (mode == MODE_V ? a.p : a.s);
      const double2* opbase = mode == MODE_V ? a.rp : (mode == MODE_T ? a.s : nullptr);
      const double* ccs = mode == MODE_RHSP ? S.cb : S.cA;
      unsigned char* my = cb_stage + (size_t)warp * 2 * CB_STAGE;
      // item k of this warp: slice, member, extent.  width < 0: no such item
      auto load_item = [&](int k, int& m, int& s, int& ai, int& base, int& width) {
        const int v = NWB * k + ((warp + 7 * k) & (NWB - 1));
        width = -1;
        m = s = ai = base = 0;
        if (v < nitems) {
          const int sl = v / nact;
          ai = v - sl * nact;
          s = slo + sl;
          m = S.act[ai];
          base = __ldg(a.slice_ptr + s);
          width = (__ldg(a.slice_ptr + s + 1) - base) >> 5;
        }
      };
      auto issue = [&](int m, int base, int width, int j0, int b) {   // columns j0 .. j0 + CB_CH - 1 of a slice -> stage b
        const int32_t* cp = a.sell_col + base + lane + j0 * 32;
        const double2* vp = Vb + (size_t)sdir[m] * a.mat_stride_sell + base + lane + j0 * 32;
        const uint32_t sc = smem_u32(my + b * CB_STAGE) + lane * 4, sv = smem_u32(my + b * CB_STAGE + CB_CH * 128) + lane * 16;
#pragma unroll
        for (int u = 0; u < CB_CH; ++u)
          if (j0 + u < width) {
            asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(sc + u * 128), "l"(cp + u * 32), "l"(l2_first) : "memory");
            asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sv + u * 512), "l"(vp + u * 32), "l"(l2_first) : "memory");
          }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      int cm, cs, cai, cbase, cwidth, nm, ns, nai, nbase, nwidth;
      load_item(0, cm, cs, cai, cbase, cwidth);
      load_item(1, nm, ns, nai, nbase, nwidth);
      int k = 1, pb = 0;
      if (cwidth >= 0) issue(cm, cbase, cwidth, 0, 0);
      while (cwidth >= 0) {
        const int row = __ldg(a.sell_row + cs * 32 + lane);
        const size_t off = (size_t)cm * vs + (row >= 0 ? row : 0);
        double2 op = make_double2(0.0, 0.0);
        if (opbase && row >= 0) op = opbase[off];
        const double2* x = xbase + (size_t)cm * vs;
        const double cc = ccs[cm];
        double ar = 0.0, ai_ = 0.0;
        int j0 = 0;
        do {
          // the piece after this one: further columns of this slice, or the first ones of the next item
          if (j0 + CB_CH < cwidth) issue(cm, cbase, cwidth, j0 + CB_CH, pb ^ 1);
          else if (nwidth >= 0) issue(nm, nbase, nwidth, 0, pb ^ 1);
          else asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 1;" ::: "memory");   // this piece has landed (every lane reads only what it copied)
          const int32_t* sc = reinterpret_cast<const int32_t*>(my + pb * CB_STAGE) + lane;
          const double2* sv = reinterpret_cast<const double2*>(my + pb * CB_STAGE + CB_CH * 128) + lane;
          double2 xv[CB_CH];
#pragma unroll
          for (int u = 0; u < CB_CH; ++u)
            if (j0 + u < cwidth) xv[u] = ldv_gather_f64x2(x + sc[u * 32]);
#pragma unroll
          for (int u = 0; u < CB_CH; ++u)
            if (j0 + u < cwidth) {
              const double2 val = sv[u * 32];
              const double pa = val.x, pb_ = cc * val.y;
              ar = fma(pa, xv[u].x, ar);
              ar = fma(-pb_, xv[u].y, ar);
              ai_ = fma(pa, xv[u].y, ai_);
              ai_ = fma(pb_, xv[u].x, ai_);
            }
          pb ^= 1;
          j0 += CB_CH;
        } while (j0 < cwidth);
        double t0 = 0.0, t1 = 0.0;
        if (row >= 0) {
          const double2 y = make_double2(ar, ai_);
          if (mode == MODE_V) {
            a.v[off] = y;
            t0 = y.x * op.x + y.y * op.y;
          } else if (mode == MODE_T) {
            a.t[off] = y;
            t0 = op.x * y.x + op.y * y.y;
            t1 = y.x * y.x + y.y * y.y;
          } else {
            a.r[off] = y;
            a.rp[off] = y;
            t0 = y.x * y.x + y.y * y.y;
          }
        }
        t0 = warp_sum(t0);
        if (mode == MODE_T) t1 = warp_sum(t1);
        if (lane == 0) { wm[warp][cm][0] += t0; wm[warp][cm][1] += t1; }
        cm = nm; cs = ns; cai = nai; cbase = nbase; cwidth = nwidth;
        ++k;
        load_item(k, nm, ns, nai, nbase, nwidth);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (a.prof && threadIdx.x == 0) a.prof[16 + blockIdx.x] += global_ns() - tp0;   // per-block time inside the passes
    PROF(mode == MODE_RHSP ? 0 : (mode == MODE_V ? 3 : 7));
    gs.target = S.bar_target;
    if (mode == MODE_V) {
      // ---- alpha = rho / (r^, v) ; s = r - alpha v
      cb_grid_reduce<NT>(a, S, wacc, acc0, acc1, ai0, 1, 2, gs, wm);
      PROF(4);
      if ((int)threadIdx.x < nact) {
        const int m = S.act[threadIdx.x];
        const double d = S.tot[2 * m];
        if (d == 0.0) { S.reason[m] = BTFEM_EBREAKDOWN; S.its[m] = S.it; atomicMin(&S.fail, (int)BTFEM_EBREAKDOWN); }
        else S.alpha[m] = S.rho[m] / d;
      }
      __syncthreads();
      if (!S.fail) {
        for (long long i = clo + threadIdx.x; i < chi; i += NT) {
          const int ai = (int)(i / n), row = (int)(i - (long long)ai * n), m = S.act[ai];
          const size_t off = (size_t)m * vs + row;
          const double alpha = S.alpha[m];
          const double2 rr = a.r[off], vv = a.v[off];
          a.s[off] = make_double2(rr.x - alpha * vv.x, rr.y - alpha * vv.y);
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
      cb_grid_reduce<NT>(a, S, wacc, acc0, acc1, ai0, 2, 3, gs, wm);
      PROF(8);
      if ((int)threadIdx.x < nact) {
        const int m = S.act[threadIdx.x];
        S.omega[m] = (S.tot[2 * m + 1] == 0.0) ? 0.0 : S.tot[2 * m] / S.tot[2 * m + 1];
      }
      __syncthreads();
      acc0[0] = acc0[1] = acc1[0] = acc1[1] = 0.0;
      {
        const bool fresh = S.it == 0;   // zero initial guess: x starts from 0
        for (long long i = clo + threadIdx.x; i < chi; i += NT) {
          const int ai = (int)(i / n), row = (int)(i - (long long)ai * n), m = S.act[ai];
          const size_t off = (size_t)m * vs + row;
          const double alpha = S.alpha[m], omega = S.omega[m];
          const double2 pp = a.p[off], ss = a.s[off], tt = a.t[off], qq = a.rp[off];
          double2 xx = make_double2(0.0, 0.0);
          if (!fresh) xx = a.u[off];
          xx.x += alpha * pp.x + omega * ss.x;
          xx.y += alpha * pp.y + omega * ss.y;
          a.u[off] = xx;
          const double2 rr = make_double2(ss.x - omega * tt.x, ss.y - omega * tt.y);
          a.r[off] = rr;
          const double t0 = rr.x * qq.x + rr.y * qq.y, t1 = rr.x * rr.x + rr.y * rr.y;
          if (ai == aiv0) { acc0[0] += t0; acc0[1] += t1; }
          else { acc1[0] += t0; acc1[1] += t1; }
        }
      }
      PROF(9);
      cb_grid_reduce<NT>(a, S, wacc, acc0, acc1, aiv0, 2, 5, gs);
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
          else if (rnorm <= S.ttol[m]) reason = rnorm < a.ctrl->atol ? 3 : 2;
          else if (rnorm >= a.ctrl->dtol * S.bn[m]) reason = BTFEM_EDTOL;
          else if (rho_used == 0.0 || omega == 0.0) reason = BTFEM_EBREAKDOWN;
          else if (it >= a.ctrl->maxit) reason = BTFEM_ENOTCONV;
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
          if (S.active != before) cb_rebuild(S, M);
        }
      }
      __syncthreads();
    } else {
      // ---- ||r|| of every member: start of the Krylov solves of this step
      cb_grid_reduce<NT>(a, S, wacc, acc0, acc1, ai0, 1, 0, gs, wm);
      PROF(10);
      if (threadIdx.x < 32) {
        const unsigned int before = S.active;
        __syncwarp();
        if (lane < nact) {
          const int m = S.act[lane];
          const double atol = a.ctrl->atol;
          const double bn = sqrt(S.tot[2 * m]);
          const double ttol = fmax(a.ctrl->rtol * bn, atol);
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
          if (S.active != before) cb_rebuild(S, M);
        }
      }
      __syncthreads();
    }
    if (!S.fail && S.active != 0u) {
      // ---- p <- r - omega*beta*v + beta*p   (first iteration: p = r), then v = A p.  The active set may just have
      // shrunk: the chunk of this block is cut again from the members that go on
      const int nact2 = S.nact;
      const long long WV2 = (long long)nact2 * n;
      const long long plo = WV2 * blockIdx.x / NB, phi = WV2 * (blockIdx.x + 1) / NB;
      const bool first = S.it == 0;
      for (long long i = plo + threadIdx.x; i < phi; i += NT) {
        const int ai = (int)(i / n), row = (int)(i - (long long)ai * n), m = S.act[ai];
        const size_t off = (size_t)m * vs + row;
        const double2 rr = a.r[off];
        if (first) {
          a.p[off] = rr;
        } else {
          const double beta = S.beta[m], ob = S.omega[m] * beta;
          const double2 vv = a.v[off];
          double2 pp = a.p[off];
          pp.x = rr.x - ob * vv.x + beta * pp.x;
          pp.y = rr.y - ob * vv.y + beta * pp.y;
          a.p[off] = pp;
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
          double2* u = a.u + (size_t)m * vs;
          for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += NB * NT) u[i] = make_double2(0.0, 0.0);
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
          KrylovCtrl* ctrl = a.ctrl + m;
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
        cb_rebuild(S, M);
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.gridbar[32] = S.bar_target;
#undef PROF
}
