// Internal declarations of libbtfem.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/btfem.h"

// SM count of the current device (148 on B200: 2 dies x 74 SMs), read once per device from the runtime
inline int bt_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    cached[dev] = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
  }
  return cached[dev];
}
#define BT_NUM_SMS bt_num_sms()
#define BT_MAX_PARTIALS 4096      // upper bound on the grid of any reducing kernel
#define BT_SELL_SIGMA 1024        // sorting window (rows) of the SELL-32 layout
// warp-stream layout of the operator for the TMA kernels (solve.cu): every warp of the launch owns a contiguous
// byte stream of "pieces" (<= PS_W columns of one SELL slice: w x 32 column indices, then w x 32 value pairs, and --
// behind the last piece of a slice -- the 32 row numbers), which it pulls through a ring of shared-memory stages with
// cp.async.bulk + mbarrier.  Column indices are 16-bit offsets from a per-piece reference column (18 bytes per
// nonzero) when every piece spans fewer than 65 536 columns -- the passes are bandwidth-bound -- else int32 (20 bytes).
// Offsets are counted in units of 64 bytes.
#define BT_PS_MAX_WARPS 16      // warps per block of the stream kernels: 8 (ring depth 4), 12 (3) or 16 (2)
#define BT_PS_W 8               // columns per piece
#define BT_PS_STAGE ((BT_PS_W * 10 + 2) * 64)   // bytes of a ring stage: a full int32 piece + the row block
// 64-byte units per stream column: 32 lanes x (2 | 4 bytes of column + 16 bytes of values)
__host__ __device__ inline int bt_ps_colu(bool c16) { return c16 ? 9 : 10; }
// byte offset of column j (< width) of a slice whose stream starts at unit u0: piece p = j / W starts at u0 + p W colu
__host__ __device__ inline size_t bt_ps_col_off(int64_t u0, int j, int lane, bool c16) {
  const int p = j / BT_PS_W;
  return (size_t)(u0 + (int64_t)p * BT_PS_W * bt_ps_colu(c16)) * 64 + (size_t)(j - p * BT_PS_W) * (c16 ? 64 : 128) +
         (size_t)lane * (c16 ? 2 : 4);
}
__host__ __device__ inline size_t bt_ps_val_off(int64_t u0, int width, int j, int lane, bool c16) {
  const int p = j / BT_PS_W;
  const int w = width - p * BT_PS_W < BT_PS_W ? width - p * BT_PS_W : BT_PS_W;
  return (size_t)(u0 + (int64_t)p * BT_PS_W * bt_ps_colu(c16)) * 64 + (size_t)w * (c16 ? 64 : 128) +
         (size_t)(j - p * BT_PS_W) * 512 + (size_t)lane * 16;
}
__host__ __device__ inline size_t bt_ps_row_off(int64_t u0, int width, int lane, bool c16) {   // behind the last piece
  return (size_t)(u0 + (int64_t)(width > 0 ? width : 0) * bt_ps_colu(c16)) * 64 + (size_t)lane * 4;
}

struct BtError {
  int code;
  std::string msg;
};

#define BT_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      throw BtError{e_ == cudaErrorMemoryAllocation ? BTFEM_ENOMEM : BTFEM_ECUDA,          \
                    std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")"};                             \
    }                                                                                      \
  } while (0)

#define BT_REQUIRE(cond, text)                         \
  do {                                                 \
    if (!(cond)) throw BtError{BTFEM_EINVAL, (text)}; \
  } while (0)

// Allocation context of the calling thread: inside an API call (api.cu: guarded) device arrays come from the
// device's stream-ordered memory pool on the handle's stream (cudaMallocAsync / cudaFreeAsync; the pool keeps what
// is freed -- release threshold = max -- so re-building a problem of the same size costs no cudaMalloc/cudaFree,
// which were 0.2-0.4 s of an end-to-end 1 M-DOF solve).  Outside (handle destruction) frees are plain cudaFree.
inline thread_local cudaStream_t bt_alloc_stream = nullptr;
inline thread_local bool bt_alloc_pooled = false;

template <typename T>
struct DevArray {
  T* p = nullptr;
  size_t n = 0;
  bool pooled = false;   // p came from the stream-ordered pool
  bool plain = false;    // always cudaMalloc: memory that is exported to peers (cudaIpcGetMemHandle)
  DevArray() {}
  DevArray(const DevArray&) = delete;
  DevArray& operator=(const DevArray&) = delete;
  ~DevArray() { release(); }
  void release() {
    if (p) {
      if (pooled && bt_alloc_pooled) cudaFreeAsync(p, bt_alloc_stream);
      else cudaFree(p);   // also valid for pool memory (synchronises)
    }
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    if (count == n && p) return;
    release();
    if (count) {
      if (!plain && bt_alloc_pooled) {
        BT_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), bt_alloc_stream));
        pooled = true;
      } else {
        BT_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        pooled = false;
      }
    }
    n = count;
  }
  void upload(const T* src, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) BT_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* dst, cudaStream_t s) const {
    if (n) BT_CUDA(cudaMemcpyAsync(dst, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    BT_CUDA(cudaStreamSynchronize(s));
  }
  void zero(cudaStream_t s) {
    if (n) BT_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// Device-resident Krylov control block: every scalar of the BiCGStab recurrence lives here so
// that one iteration is a fixed sequence of kernel launches with constant arguments (CUDA graph).
struct KrylovCtrl {
  double rho, rho_old, alpha, omega;
  double bnorm, ttol, rnorm;
  double rtol, atol, dtol;
  double theta_cA_scale;   // theta          (A side:   c = theta * cA[step])
  double theta_cb_scale;   // -(1 - theta)   (RHS side: c = -(1-theta) * cb[step])
  int maxit;
  int iters;
  int done;
  int reason;
  int step;        // time step the iteration kernels work on
  int step_next;   // time step the next RHS launch starts
  int nonzero_guess;
  int failed;             // sticky: a step ended with a negative reason; the remaining steps are skipped
  unsigned int ticket[8];
  // device-driven time loop (the whole step is one graph with a WHILE node): statistics kept on the device
  long long total_iters;
  int max_iters;
  unsigned int member_ticket;
};

// ---- row partition of ONE mesh over several GPUs (one process per GPU, peer memory over NVLink) -------------
#define BT_MAX_RANKS 8
#define BT_COMM_ELEMS 64        // double2 elements reserved behind the vector slab for the DistComm block

// Lives in the IPC-exported allocation of every rank; written by the PEERS with system-scope stores.
// Everything that crosses NVLink uses the "LL" encoding: a double travels as two 8-byte words (sequence number
// << 32 | half of the bits) stored with one 16-byte store; 8-byte stores are single-copy atomic, so a reader that
// sees the expected sequence number in both words has the value -- no fence, no separate flag, one NVLink flight
// (measured on 2 x B200, scripts/p2p_latency.cu: 1.0 us per exchange against 4.1 us for payload +
// __threadfence_system + flag, 7-11 us for a fenced bulk push).
struct DistComm {
  unsigned long long ar_ll[2][BT_MAX_RANKS][4][2];   // all-reduce: [buffer][sender][value][half]
};

static_assert(sizeof(DistComm) <= BT_COMM_ELEMS * 16, "DistComm does not fit its reservation");

// Mutable state of a partitioned handle (device memory of the rank itself).
struct DistDev {
  unsigned long long gen[3];         // generation of the halo entries of u, p, s last published to the peers
  unsigned long long ar_seq;         // all-reduces done (never reset: every rank runs the same sequence)
  unsigned int tick[3];              // block tickets of the kernels that produce u (push), p, s
  int error;                         // a wait timed out: the solve is abandoned on every rank
  // optional timeline (btfem_dist_trace): 8 words per kernel that closes a collective
  unsigned long long* trace;
  unsigned int trace_cap, trace_pos;
};

// per-peer addresses (device memory, read-only after btfem_dist_connect; indexed by a run-time rank)
struct PeerTab {
  unsigned long long* ll[BT_MAX_RANKS];   // LL halo buffer of every rank (own or peer-mapped)
  DistComm* comm[BT_MAX_RANKS];           // comm block of every rank
  int n_ll[BT_MAX_RANKS];                 // LL entries per vector at every rank
};

// Immutable description of the partition, passed BY VALUE with the kernel arguments (scalars and pointers only,
// so that it stays in the constant bank).
struct DistView {
  int on;                            // 0: whole-mesh handle
  int rank, world;
  int n_int;                         // owned rows [0,n_int) are not needed by any peer and reference no halo column
  int halo_begin;                    // vector element index of the first halo dof (n_own rounded up to 8)
  int n_ll;                          // LL entries per vector here: halo dofs, then periodic-gather sources
  int wait_slice;                    // SELL slices below this one never touch a halo column
  int n_send_u;                      // entries of the per-step u push (halo dofs + periodic sources of the peers)
  DistDev* st;
  unsigned long long* ll;            // this rank's LL buffer [3][n_ll][4 words]
  DistComm* comm;                    // this rank's comm block
  const PeerTab* peers;
  const int32_t* send_src;           // [n_send_u] local owned dof
  const int32_t* send_rank;          // [n_send_u] destination rank
  const int32_t* send_slot;          // [n_send_u] LL entry at the destination
  // the Krylov entries grouped by source row, for the update kernels that push while they produce:
  // boundary row n_int + j sends entries [bsend_ptr[j], bsend_ptr[j+1])
  const int32_t* bsend_ptr;
  const int32_t* bsend_rank;
  const int32_t* bsend_slot;
  unsigned long long timeout_ns;
};

// what the ranks exchange (through the host layer) before btfem_dist_connect
struct DistBlob {
  uint64_t magic;
  int64_t pid;
  uint64_t raw_ptr;
  int64_t device;
  int64_t npad, n_own, ndof, n_extra, halo_shift;
  cudaIpcMemHandle_t ipc;            // 64 bytes
  char pad_[BTFEM_DIST_BLOB_BYTES - 9 * 8 - 64];
};
static_assert(sizeof(DistBlob) == BTFEM_DIST_BLOB_BYTES, "DistBlob size");

struct FacetKey {
  uint32_t a, b, c;   // sorted vertex ids
  uint32_t cf;        // cell*4 + local facet
};

struct btfem {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool pool_ok = false;   // device arrays of this handle come from the stream-ordered pool (btfem_create)
  std::string err;

  // ---- inputs (host copies are kept: they are small next to the matrices and make setters order-free)
  int64_t nv = 0, nc = 0;
  int cell_nv = 4;   // vertices per cell: 4 tetrahedra, 3 triangles, 2 segments (stored in 4 slots, unused = -1)
  bool two_comp = false;
  std::vector<double> h_xyz;
  double bbox_lo[3] = {0, 0, 0}, bbox_hi[3] = {0, 0, 0};   // bounding box of the vertices (btfem_get_bbox)
  int dkind = 0;
  std::vector<double> h_D{1.0};
  int t2kind = 0;
  std::vector<double> h_invT2{0.0};
  int kkind = 0;
  std::vector<double> h_kappa{0.0};
  int32_t nmark = 0;
  std::vector<int32_t> h_marker;
  bool periodic = false;
  int32_t pdir[3] = {0, 0, 0};
  double kappa_e = 0.0, ptol = 0.0, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  std::vector<double> h_ic;   // empty = all ones
  // strongly imposed periodicity (btfem_set_periodic_map): master vertex of every vertex; empty = none.
  // Non-empty also selects the transformed equation (strong.cu) in btfem_solve.
  std::vector<int32_t> h_vmaster;

  // ---- device mesh
  DevArray<double> d_xyz, d_D, d_invT2, d_kappa_tab, d_bmark /*kappa_e^h per vertex*/;
  DevArray<int32_t> d_tets, d_phase, d_marker, d_cell_dofs, d_dof_vertex, d_dof_comp, d_vc2dof;
  std::vector<int32_t> h_dof_vertex, h_dof_comp;

  // ---- facets
  int64_t n_iface = 0, n_bfacet = 0;
  DevArray<int32_t> d_if_dofs;     // [n_iface*6]: d0[3], d1[3]
  DevArray<int32_t> d_if_verts;    // [n_iface*3]
  DevArray<double> d_if_kappa;     // [n_iface]
  DevArray<int32_t> d_bf_verts;    // [n_bfacet*3]
  DevArray<int32_t> d_bf_dofs;     // [n_bfacet*3]

  // ---- row partition (nv_own < 0: the handle owns the whole mesh)
  int64_t nv_own = -1, nv_int = -1;   // vertices [0,nv_own) owned, of which [0,nv_int) are not needed by peers
  int64_t n_own = 0, n_int = 0;       // the same in dofs (set by the dof map)
  int64_t halo_shift = 0;             // halo dof j sits at vector element j + halo_shift (128-byte aligned halo)
  int64_t n_extra = 0;                // u values of periodic-gather sources owned by peers, behind the DistComm block
  bool dist_connected = false, dist_failed = false;
  int rank = 0, world = 1;
  DevArray<DistDev> d_dist;
  DistView dview{};                   // filled by btfem_dist_connect
  DevArray<PeerTab> d_peers;
  DevArray<unsigned long long> d_trace;
  int64_t trace_cap = 0;
  DevArray<int32_t> d_send_src, d_send_rank, d_send_slot, d_bsend_ptr, d_bsend_rank, d_bsend_slot;
  void* peer_map[BT_MAX_RANKS] = {nullptr};   // cudaIpcOpenMemHandle mappings to close
  int64_t n_rows() const { return nv_own >= 0 ? n_own : ndof; }

  // ---- pattern
  bool assembled = false;
  int64_t ndof = 0, nnz = 0, nsrc = 0;
  DevArray<int32_t> d_rowptr, d_colidx, d_rowidx, d_diagpos;
  // SELL-32 copy of the pattern for the fused SpMV (rows sorted by length inside windows of BT_SELL_SIGMA)
  int64_t n_slice = 0, nnz_sell = 0;
  DevArray<int32_t> d_slice_ptr;   // [n_slice+1]
  // static warp schedule of the fused SpMV (setup.cu: list scheduling of the slices over the warps of the launch)
  DevArray<int32_t> d_sched;       // [n_slice] slice ids, grouped by warp
  DevArray<int32_t> d_sched_ptr;   // [2*sched_warps+1]: warp w runs sched[ptr[2w] .. ptr[2w+1]) with halo reads,
                                   //                    then sched[ptr[2w+1] .. ptr[2w+2]) plain
  int sched_grid = 0;              // blocks of the launch the schedule was built for
  DevArray<int32_t> d_sell_row;    // [n_slice*32] slot -> row (-1 = padding slot)
  DevArray<int32_t> d_sell_slot;   // [ndof] row -> slot
  DevArray<int32_t> d_sell_col;    // [nnz_sell]
  DevArray<double2> d_PJs, d_QJs;  // [nnz_sell] per-solve operator values in SELL order
  // batched solves (btfem_solve_batch): ONE direction-independent copy of the operator serves all members --
  // (P_k, Q_k)/P_rr, (Jx_k, Jy_k), Jz_k in SELL order; J_g is formed per member inside the SpMV
  DevArray<double2> d_PQs, d_Jxys;
  DevArray<double> d_Jzs, d_gdirs /*[members*3]*/;
  // warp-stream copy of the SELL operator (whole-mesh handles; built with the pattern, filled by bt_combine)
  int ps_req_blocks = 0;           // btfem_set_sm_partition: blocks the stream kernels may use (0 = all SMs)
  int ps_blocks = 0;               // blocks (= SMs) the layout was built for; 0 = none
  int ps_warps = 8;                // warps per block the layout was built for
  int64_t ps_units = 0;            // length of a stream in 64-byte units
  bool ps_col16 = false;           // 16-bit column offsets from a per-piece reference column
  int ps_max_pieces = 0;           // longest piece list of a warp
  DevArray<int32_t> d_ps_ptr;      // [ps_blocks * ps_warps + 1] piece range of every warp
  DevArray<int4> d_ps_piece;       // {stream offset (units), columns, slice, reference column << 1 | last piece of its slice}
  DevArray<int32_t> d_ps_scol0;    // [n_slice] stream offset (units) of the first piece of the slice
  DevArray<int32_t> d_ps_first;    // [n_slice] index of the slice's first piece in d_ps_piece
  DevArray<unsigned char> d_PJt, d_QJt;   // [ps_units * 64] per-solve operator values + columns + rows, stream order
  DevArray<int32_t> d_member_dir;          // batch: operator copy used by every member
  DevArray<long long> d_cb_cost;           // many-warp batch kernel: L1-wavefront cost prefix over the slices
  DevArray<unsigned char> d_PJt_b, d_QJt_b;   // persistent batch kernel: one stream pair per gradient direction of the batch
  DevArray<uint32_t> d_src;        // contribution ids sorted by (row,col), stable
  DevArray<int64_t> d_seg;         // [nnz+1] segment offsets into d_src
  DevArray<double> d_vals[8];      // M,S,R,Jx,Jy,Jz,I,B
  DevArray<double> d_lumped, d_ic_dof;
  double whole_vol = 0, voi = 0, voi_comp[2] = {0, 0};

  // ---- per-solve state
  DevArray<double2> d_PJ, d_QJ;
  DevArray<double> d_Bhat, d_dinv;
  DevArray<double> d_strongW, d_strongG;   // strong periodic BC: W and G = C - N of the current direction (strong.cu)
  // the seven Krylov vectors live in ONE slab so that a single L2 access-policy window can pin them:
  // 7 x 8 MB at 1 M DOFs fits the 126 MB L2 while the matrix streams past with evict-first loads
  DevArray<double2> d_vecs;
  struct VecView {
    double2* p = nullptr;
    size_t n = 0;
    void zero(cudaStream_t s) { if (n) BT_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double2), s)); }
    void download(double2* dst, cudaStream_t s) const {
      if (n) BT_CUDA(cudaMemcpyAsync(dst, p, n * sizeof(double2), cudaMemcpyDeviceToHost, s));
      BT_CUDA(cudaStreamSynchronize(s));
    }
  } d_u, d_r, d_rp, d_p, d_v, d_s, d_t;
  DevArray<double2> d_gm_V;        // GMRES basis, (restart+1) vectors
  DevArray<double> d_gm_h;
  DevArray<double> d_gm_state;     // sizeof(GmState) (solve.cu): device-resident Hessenberg / Givens / residual state
  double* h_gm = nullptr;          // pinned
  bool l2_window_set = false;
  cudaAccessPolicyWindow l2_window{};
  DevArray<double> d_cA, d_cb, d_Fb;
  DevArray<double> d_partials;     // [8][BT_MAX_PARTIALS]
  DevArray<KrylovCtrl> d_ctrl;
  // ILU(0) preconditioner (ilu.cu): complex factors on the CSR pattern, scratch, ready flags with an epoch
  DevArray<double2> d_ilu, d_ilu_y, d_ilu_tmp;
  DevArray<unsigned int> d_ilu_flag;
  unsigned int ilu_epoch = 0;
  double ilu_c = 0.0;
  bool ilu_valid = false;
  DevArray<unsigned int> d_gridbar;   // persistent BiCGStab kernel: [0] arrival counter, [32] its value at kernel start
  KrylovCtrl* h_ctrl = nullptr;    // pinned, one per batch member
  int h_ctrl_n = 0;
  size_t vec_npad = 0;             // padded vector length (elements)
  int64_t step_stride = 0;         // cA/cb stride between batch members
  int comb_members = 0;
  double comb_dt = -1, comb_theta = -1, comb_g[3] = {0, 0, 0};
  int comb_pc = -1;
  bool have_solution = false;
  int lanes = 0;                   // fused SpMV variant: 0 = SELL-32 (default), else CSR with `lanes` threads per row

  // weak pseudo-periodic BC: gather operator (per boundary dof: 3 sources, weights, displacement)
  int64_t n_pb = 0;
  DevArray<int32_t> d_pb_dof, d_pb_src;   // [n_pb], [n_pb*3]
  DevArray<double> d_pb_w, d_pb_dx;       // [n_pb*3], [n_pb*3]
  DevArray<int32_t> d_pb_rows;            // rows with a nonzero in B (they receive (1-theta)*B*u_bc)
  int64_t n_pb_rows = 0;
  DevArray<double2> d_ubc, d_rhs_add;     // [ndof] dense, zero off the boundary
};

// setup.cu
void bt_mesh_stats(btfem* h, double* hmin, double* hmax);
void bt_build_dofmap(btfem* h);
void bt_build_facets(btfem* h);
void bt_build_pattern(btfem* h);
void bt_assemble_values(btfem* h);
void bt_build_periodic(btfem* h);

// ilu.cu
void bt_ilu_factor(btfem* h, double c, cudaStream_t st);
void bt_ilu_apply(btfem* h, const double2* in, double2* out, cudaStream_t st);
void bt_ilu_get(btfem* h, double* out);
// solve.cu
bool bt_stream_kernel_usable(const btfem* h);   // warp-stream layout built for this device and not switched off
void bt_combine(btfem* h, double dt, double theta, const double g[3], int pc, int member = 0, int members = 1);
void bt_strong_build(btfem* h, const double g[3]);
void bt_strong_recombine(btfem* h, cudaStream_t st, double dt, double theta, int pc);
void bt_strong_get(btfem* h, const double g[3], double* W, double* G);
void bt_spmv_host(btfem* h, double dt, double theta, double c, const double g[3], const double* x, double* y);
void bt_spmv_bench(btfem* h, double dt, double theta, double c, const double g[3], int lanes, int nrep, int flush_l2,
                   double* ms);
void bt_solve(btfem* h, const btfem_solve_args* a, btfem_solve_out* out, int32_t* iters_per_step);
void bt_solve_batch(btfem* h, int members, const btfem_solve_args* a, btfem_solve_out* out);
void bt_dist_export(btfem* h, void* blob);
void bt_dist_connect(btfem* h, int rank, int world, const void* blobs, int64_t nsend, const int32_t* src,
                     const int32_t* dst_rank, const int32_t* dst_slot, int64_t nsend_u, const int32_t* src_u,
                     const int32_t* dst_rank_u, const int32_t* dst_index_u, const int32_t* recv_from);
void bt_dist_close(btfem* h);
int64_t bt_dist_get_trace(btfem* h, uint64_t* out, int64_t max_entries);
