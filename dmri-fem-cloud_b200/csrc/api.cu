// extern "C" entry points of libbtfem.so: argument checking, error capture, call order.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "btfem_internal.cuh"

namespace {

template <typename F>
int guarded(btfem* h, F&& f) {
  if (!h) return BTFEM_EINVAL;
  struct AllocScope {   // device arrays touched inside this call live on the handle's stream (btfem_internal.cuh)
    cudaStream_t s0 = bt_alloc_stream;
    bool p0 = bt_alloc_pooled;
    explicit AllocScope(btfem* h) { bt_alloc_stream = h->stream; bt_alloc_pooled = h->pool_ok; }
    ~AllocScope() { bt_alloc_stream = s0; bt_alloc_pooled = p0; }
  } scope(h);
  try {
    BT_CUDA(cudaSetDevice(h->device));
    f();
    h->err.clear();
    return BTFEM_OK;
  } catch (const BtError& e) {
    h->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return BTFEM_EINVAL;
  }
}

void invalidate(btfem* h) {
  h->assembled = false;
  h->n_pb = 0;
  h->comb_dt = -1;
  h->ilu_valid = false;
  h->have_solution = false;
  bt_dist_close(h);   // the peers' halo maps describe the old numbering
}

}  // namespace

extern "C" {

int btfem_version(void) { return 100; }

int btfem_create(int device, btfem_t** out) {
  if (!out) return BTFEM_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return BTFEM_ECUDA;
  btfem* h = new btfem();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return BTFEM_ECUDA;
  }
  // Stream-ordered pool for the handle's device arrays: freed blocks stay in the pool (no trim at synchronisation
  // points), so tearing a problem down and building the next one re-uses them.  BTFEM_POOL=0: plain cudaMalloc.
  {
    const char* env = getenv("BTFEM_POOL");
    int supported = 0;
    cudaMemPool_t pool = nullptr;
    if (!(env && env[0] == '0') &&
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && supported &&
        cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      h->pool_ok = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
    }
    cudaGetLastError();
  }
  h->d_vecs.plain = true;   // the vector slab of a row partition is exported to the peers (cudaIpcGetMemHandle)
  *out = h;
  return BTFEM_OK;
}

void btfem_destroy(btfem_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  bt_dist_close(h);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->h_gm) cudaFreeHost(h->h_gm);
  cudaStream_t st = h->stream;
  {
    // device arrays go back to the stream-ordered pool (cudaFreeAsync), not to the driver: the next handle on this
    // device -- a sweep or an end-to-end loop builds one per problem -- gets them from there instead of paying for
    // ~1.5 GB of cudaFree + cudaMalloc (0.05-0.3 s at 1 M DOFs)
    const cudaStream_t s0 = bt_alloc_stream;
    const bool p0 = bt_alloc_pooled;
    bt_alloc_stream = st;
    bt_alloc_pooled = h->pool_ok && st != nullptr;
    delete h;   // frees device arrays
    bt_alloc_stream = s0;
    bt_alloc_pooled = p0;
  }
  if (st) {
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
}

const char* btfem_last_error(btfem_t* h) { return h ? h->err.c_str() : "null handle"; }

static void set_mesh_cells(btfem_t* h, int64_t nv, const double* xyz, int64_t nc, const int32_t* cells, int cell_nv,
                           const int32_t* phase) {
  BT_REQUIRE(nv > 0 && nc > 0 && xyz && cells, "empty mesh");
  BT_REQUIRE(nv < (1LL << 30) && nc < (1LL << 27), "mesh too large for 32-bit indices");
  invalidate(h);
  h->h_vmaster.clear();
  if (nv != h->nv || nc != h->nc) {   // per-cell / per-vertex inputs of the previous mesh do not fit this one
    if (h->dkind != 0) { h->dkind = 0; h->h_D.assign(1, 1.0); }
    if (h->t2kind != 0) { h->t2kind = 0; h->h_invT2.assign(1, 0.0); }
    if (h->kkind != 0) { h->kkind = 0; h->h_kappa.assign(1, 0.0); h->nmark = 0; }
    h->h_marker.clear();
    h->h_ic.clear();
  }
  h->nv = nv;
  h->nc = nc;
  h->cell_nv = cell_nv;
  h->two_comp = phase != nullptr;
  // Range checks and the host copies run on a helper thread while this one feeds the copy engine (a 1 M-DOF mesh has
  // 11 M indices and 70 MB to upload).  Device kernels never see the mesh before the checks have passed: the handle
  // stays un-assembled and the error is raised below.
  int32_t lo = 0, hi = 0, bad_phase = 0;
  const bool timing = getenv("BTFEM_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[btfem] set_mesh: %-18s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  };
  std::thread helper([&] {
    auto h0 = std::chrono::steady_clock::now();
    for (int64_t i = 0; i < cell_nv * nc; ++i) {   // branch-free min/max passes: they vectorise
      lo = cells[i] < lo ? cells[i] : lo;
      hi = cells[i] > hi ? cells[i] : hi;
    }
    if (phase)
      for (int64_t i = 0; i < nc; ++i) bad_phase |= phase[i] & ~1;
    h->h_xyz.assign(xyz, xyz + 3 * nv);   // kept: the periodic marker is evaluated on the host (setup.cu)
    double lo3[3] = {xyz[0], xyz[1], xyz[2]}, hi3[3] = {xyz[0], xyz[1], xyz[2]};
    for (int64_t v = 0; v < nv; ++v)      // bounding box (GetGlobalDomainSize, DmriFemLib.py:560-581)
      for (int d = 0; d < 3; ++d) {
        const double x = xyz[3 * v + d];
        lo3[d] = x < lo3[d] ? x : lo3[d];
        hi3[d] = x > hi3[d] ? x : hi3[d];
      }
    for (int d = 0; d < 3; ++d) { h->bbox_lo[d] = lo3[d]; h->bbox_hi[d] = hi3[d]; }
    if (timing)
      fprintf(stderr, "[btfem] set_mesh: helper thread      %7.2f ms\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
  });
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } joiner{helper};
  h->d_xyz.upload(xyz, 3 * nv, h->stream);
  if (phase) h->d_phase.upload(phase, nc, h->stream); else h->d_phase.release();
  lap("xyz/phase enqueue");
  if (cell_nv == 4) {
    h->d_tets.upload(cells, 4 * nc, h->stream);
    lap("tets enqueue");
    helper.join();
    lap("join helper");
  } else {
    helper.join();
    std::vector<int32_t> slots(4 * nc, -1);   // triangles / segments keep the 4-slot cell layout, unused slots = -1
    for (int64_t c = 0; c < nc; ++c)
      for (int k = 0; k < cell_nv; ++k) slots[4 * c + k] = cells[cell_nv * c + k];
    h->d_tets.upload(slots.data(), 4 * nc, h->stream);
    BT_CUDA(cudaStreamSynchronize(h->stream));   // `slots` is a temporary
  }
  BT_CUDA(cudaStreamSynchronize(h->stream));
  lap("stream sync");
  if (!(lo >= 0 && hi < nv) || bad_phase) {
    h->nv = h->nc = 0;   // nothing usable was set
    BT_REQUIRE(lo >= 0 && hi < nv, "cell vertex index out of range");
    BT_REQUIRE(bad_phase == 0, "phase must be 0 or 1");
  }
}

int btfem_set_mesh(btfem_t* h, int64_t nv, const double* xyz, int64_t nc, const int32_t* tets, const int32_t* phase) {
  return guarded(h, [&] { set_mesh_cells(h, nv, xyz, nc, tets, 4, phase); });
}

int btfem_set_mesh_tri(btfem_t* h, int64_t nv, const double* xyz, int64_t nc, const int32_t* tris,
                       const int32_t* phase) {
  return guarded(h, [&] { set_mesh_cells(h, nv, xyz, nc, tris, 3, phase); });
}

int btfem_set_mesh_seg(btfem_t* h, int64_t nv, const double* xyz, int64_t nc, const int32_t* segs) {
  return guarded(h, [&] { set_mesh_cells(h, nv, xyz, nc, segs, 2, nullptr); });
}

int btfem_set_periodic_map(btfem_t* h, const int32_t* vmaster) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nv > 0, "set the mesh first");
    BT_REQUIRE(h->nv_own < 0, "strong periodic BC: whole-mesh handles");
    if (vmaster) {
      BT_REQUIRE(h->cell_nv >= 3, "strong periodic BC: tetrahedral or triangle meshes");
      for (int64_t v = 0; v < h->nv; ++v) {
        BT_REQUIRE(vmaster[v] >= 0 && vmaster[v] < h->nv, "periodic map: master out of range");
        BT_REQUIRE(vmaster[vmaster[v]] == vmaster[v], "periodic map: the master of a master must be itself");
      }
      h->h_vmaster.assign(vmaster, vmaster + h->nv);
    } else {
      h->h_vmaster.clear();
    }
    invalidate(h);
  });
}

int btfem_get_strong_operators(btfem_t* h, const double gdir[3], double* W, double* G) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && gdir, "call btfem_assemble first");
    bt_strong_get(h, gdir, W, G);
  });
}

int btfem_set_phase(btfem_t* h, const int32_t* phase) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nc > 0, "set the mesh first");
    BT_REQUIRE(!phase || h->cell_nv > 2, "segment meshes are one-compartment");
    if (phase)
      for (int64_t i = 0; i < h->nc; ++i) BT_REQUIRE(phase[i] == 0 || phase[i] == 1, "phase must be 0 or 1");
    invalidate(h);
    h->two_comp = phase != nullptr;
    if (phase) {
      h->d_phase.upload(phase, h->nc, h->stream);
      BT_CUDA(cudaStreamSynchronize(h->stream));
    } else {
      h->d_phase.release();
    }
  });
}

int btfem_get_mesh_stats(btfem_t* h, double* hmin, double* hmax) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nc > 0 && hmin && hmax, "set the mesh first");
    bt_mesh_stats(h, hmin, hmax);
  });
}

int btfem_get_bbox(btfem_t* h, double lo[3], double hi[3]) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nv > 0 && lo && hi, "set the mesh first");
    for (int d = 0; d < 3; ++d) { lo[d] = h->bbox_lo[d]; hi[d] = h->bbox_hi[d]; }
  });
}

int btfem_set_diffusion(btfem_t* h, int kind, const double* D) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nc > 0, "set the mesh first");
    BT_REQUIRE(D && kind >= 0 && kind <= 2, "bad diffusion kind");
    size_t n = kind == 0 ? 1 : (kind == 1 ? (size_t)h->nc : (size_t)9 * h->nc);
    h->dkind = kind;
    h->h_D.assign(D, D + n);
    invalidate(h);
  });
}

int btfem_set_relaxation(btfem_t* h, int kind, const double* inv_t2) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nc > 0, "set the mesh first");
    BT_REQUIRE(inv_t2 && kind >= 0 && kind <= 1, "bad relaxation kind");
    h->t2kind = kind;
    h->h_invT2.assign(inv_t2, inv_t2 + (kind == 0 ? 1 : (size_t)h->nc));
    invalidate(h);
  });
}

int btfem_set_permeability(btfem_t* h, int kind, const double* kappa, int32_t nmark, const int32_t* marker) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nc > 0, "set the mesh first");
    BT_REQUIRE(kappa && kind >= 0 && kind <= 1, "bad permeability kind");
    if (kind == 0) {
      h->h_kappa.assign(kappa, kappa + 1);
      h->h_marker.clear();
      h->nmark = 0;
    } else {
      BT_REQUIRE(nmark > 0 && marker, "permeability table needs cell markers");
      for (int64_t i = 0; i < h->nc; ++i) BT_REQUIRE(marker[i] >= 0 && marker[i] < nmark, "cell marker out of range");
      h->h_kappa.assign(kappa, kappa + (size_t)nmark * nmark);
      h->h_marker.assign(marker, marker + h->nc);
      h->nmark = nmark;
    }
    h->kkind = kind;
    invalidate(h);
  });
}

int btfem_set_periodic(btfem_t* h, const int32_t pdir[3], double kappa_e, double tol, const double lo[3],
                       const double hi[3]) {
  return guarded(h, [&] {
    BT_REQUIRE(pdir && lo && hi, "null argument");
    h->periodic = (pdir[0] + pdir[1] + pdir[2]) > 0;
    for (int d = 0; d < 3; ++d) {
      h->pdir[d] = pdir[d];
      h->lo[d] = lo[d];
      h->hi[d] = hi[d];
    }
    h->kappa_e = kappa_e;
    h->ptol = tol;
    invalidate(h);
  });
}

int btfem_get_boundary_facets(btfem_t* h, int32_t* verts) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && (verts || h->n_bfacet == 0), "call btfem_assemble first");
    if (h->n_bfacet) h->d_bf_verts.download(verts, h->stream);
  });
}

int btfem_set_periodic_gather(btfem_t* h, int64_t nb, const int32_t* dof, const int32_t* src, const double* w,
                              const double* dx) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled, "call btfem_assemble first");
    BT_REQUIRE(nb >= 0 && (nb == 0 || (dof && src && w && dx)), "null argument");
    const bool part = h->nv_own >= 0;
    BT_REQUIRE(!h->dist_connected, "periodic gather of a partition must be set before btfem_dist_export");
    int64_t n_extra = 0;
    for (int64_t i = 0; i < nb; ++i) {
      BT_REQUIRE(dof[i] >= 0 && dof[i] < h->ndof, "periodic gather: dof out of range");
      for (int k = 0; k < 3; ++k) {
        const int32_t s = src[3 * i + k];
        BT_REQUIRE(s < h->ndof && (part || s >= -1), "periodic gather: source out of range");
        if (s <= -2) n_extra = std::max<int64_t>(n_extra, (int64_t)(-2 - s) + 1);
      }
    }
    // sources become ELEMENT indices relative to u.  Partitioned handle: halo dofs sit halo_shift elements
    // further, and everything at or behind halo_begin = n_own + halo_shift is read from the LL buffer of u --
    // the halo dofs first, then the mirrored sources the owning peers deliver (src = -2-k -> entry n_halo + k)
    std::vector<int32_t> elem(src, src + 3 * nb);
    if (part) {
      const int64_t halo_begin = h->n_own + h->halo_shift, n_halo = h->ndof - h->n_own;
      BT_REQUIRE(halo_begin + n_halo + n_extra < (int64_t)0x7fffffffLL, "periodic source buffer exceeds int32 indexing");
      for (auto& s : elem) {
        if (s <= -2) s = (int32_t)(halo_begin + n_halo + (-2 - s));
        else if (s >= h->n_own) s += (int32_t)h->halo_shift;
      }
      if (n_extra != h->n_extra) h->d_vecs.release();
      h->n_extra = n_extra;
    }
    h->n_pb = nb;
    h->d_pb_dof.upload(dof, nb, h->stream);
    h->d_pb_src.upload(elem.data(), 3 * nb, h->stream);
    h->d_pb_w.upload(w, 3 * nb, h->stream);
    h->d_pb_dx.upload(dx, 3 * nb, h->stream);
    BT_CUDA(cudaStreamSynchronize(h->stream));
  });
}

int btfem_set_initial(btfem_t* h, const double* ic) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nv > 0, "set the mesh first");
    if (ic) h->h_ic.assign(ic, ic + h->nv); else h->h_ic.clear();
    invalidate(h);
  });
}

int btfem_assemble(btfem_t* h) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nc > 0, "set the mesh first");
    if (h->assembled) return;
    const bool timing = getenv("BTFEM_TIMING") != nullptr;   // wall clock per stage (each ends with a stream sync)
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!timing) return;
      cudaStreamSynchronize(h->stream);
      auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "[btfem] assemble: %-10s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
      t0 = t1;
    };
    bt_build_dofmap(h);
    lap("dofmap");
    bt_build_facets(h);
    lap("facets");
    bt_build_pattern(h);
    lap("pattern");
    bt_assemble_values(h);
    lap("values");
    bt_build_periodic(h);
    lap("periodic");
    h->assembled = true;
  });
}

int btfem_get_sizes(btfem_t* h, int64_t* ndof, int64_t* nnz, int64_t* n_iface, int64_t* n_bfacet) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled, "call btfem_assemble first");
    if (ndof) *ndof = h->ndof;
    if (nnz) *nnz = h->nnz;
    if (n_iface) *n_iface = h->n_iface;
    if (n_bfacet) *n_bfacet = h->n_bfacet;
  });
}

int btfem_get_pattern(btfem_t* h, int32_t* rowptr, int32_t* colidx) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && rowptr && colidx, "call btfem_assemble first");
    h->d_rowptr.download(rowptr, h->stream);
    h->d_colidx.download(colidx, h->stream);
  });
}

int btfem_get_dofmap(btfem_t* h, int32_t* dof_vertex, int32_t* dof_comp) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && dof_vertex && dof_comp, "call btfem_assemble first");
    memcpy(dof_vertex, h->h_dof_vertex.data(), sizeof(int32_t) * h->ndof);
    memcpy(dof_comp, h->h_dof_comp.data(), sizeof(int32_t) * h->ndof);
  });
}

int btfem_get_values(btfem_t* h, int which, double* out) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && out, "call btfem_assemble first");
    BT_REQUIRE(which >= 0 && which < 8, "bad matrix id");
    h->d_vals[which].download(out, h->stream);
  });
}

int btfem_get_lumped_mass(btfem_t* h, double* out) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && out, "call btfem_assemble first");
    h->d_lumped.download(out, h->stream);
  });
}

int btfem_set_lanes(btfem_t* h, int32_t lanes) {
  return guarded(h, [&] {
    BT_REQUIRE(lanes == 0 || lanes == 4 || lanes == 8 || lanes == 16 || lanes == 32,
               "lanes must be 0 (SELL-32), 4, 8, 16 or 32");
    h->lanes = lanes;
  });
}

int btfem_get_spmv_kernel(btfem_t* h, int32_t* kind) {
  return guarded(h, [&] {
    BT_REQUIRE(kind, "null argument");
    BT_REQUIRE(h->assembled, "call btfem_assemble first");
    *kind = h->lanes != 0 ? 0 : (bt_stream_kernel_usable(h) ? 2 : 1);
  });
}

int btfem_set_sm_partition(btfem_t* h, int32_t nblocks) {
  return guarded(h, [&] {
    BT_REQUIRE(nblocks >= 0 && nblocks <= BT_NUM_SMS, "SM partition: 0 (all) .. number of SMs of the device");
    if (nblocks != h->ps_req_blocks) {
      h->ps_req_blocks = nblocks;
      invalidate(h);   // the warp-stream layout belongs to one launch shape
    }
  });
}

int btfem_get_ilu_factors(btfem_t* h, double* out) {
  return guarded(h, [&] {
    BT_REQUIRE(out, "null argument");
    bt_ilu_get(h, out);
  });
}

int btfem_spmv(btfem_t* h, double dt, double theta, double c, const double gdir[3], const double* x, double* y) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && gdir && x && y, "call btfem_assemble first");
    bt_spmv_host(h, dt, theta, c, gdir, x, y);
  });
}

int btfem_spmv_bench(btfem_t* h, double dt, double theta, double c, const double gdir[3], int32_t lanes, int32_t nrep,
                     int32_t flush_l2, double* ms_per_launch) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && gdir && ms_per_launch && nrep > 0, "call btfem_assemble first");
    bt_spmv_bench(h, dt, theta, c, gdir, lanes, nrep, flush_l2, ms_per_launch);
  });
}

int btfem_solve(btfem_t* h, const btfem_solve_args* args, btfem_solve_out* out, int32_t* iters_per_step) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled, "call btfem_assemble first");
    BT_REQUIRE(args && out && args->cA && args->cb, "null argument");
    memset(out, 0, sizeof(*out));
    bt_solve(h, args, out, iters_per_step);
  });
}

int btfem_solve_batch(btfem_t* h, int32_t members, const btfem_solve_args* args, btfem_solve_out* out) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled, "call btfem_assemble first");
    BT_REQUIRE(args && out && members >= 1, "null argument");
    for (int b = 0; b < members; ++b) BT_REQUIRE(args[b].cA && args[b].cb, "null cA/cb");
    memset(out, 0, sizeof(*out) * members);
    bt_solve_batch(h, members, args, out);
  });
}

int btfem_get_solution(btfem_t* h, double* u) {
  return guarded(h, [&] {
    BT_REQUIRE(h->have_solution && u, "no solution yet");
    if (h->nv_own >= 0) memset(u, 0, sizeof(double) * 2 * h->ndof);   // halo dofs belong to peers: reported as 0
    h->d_u.download(reinterpret_cast<double2*>(u), h->stream);
  });
}

int btfem_set_partition(btfem_t* h, int64_t nv_own, int64_t nv_interior) {
  return guarded(h, [&] {
    BT_REQUIRE(h->nv > 0, "set the mesh first");
    BT_REQUIRE(h->cell_nv == 4, "row partitions are built on tetrahedral meshes");
    BT_REQUIRE(h->h_vmaster.empty(), "strong periodic BC: whole-mesh handles");
    BT_REQUIRE(nv_own > 0 && nv_own <= h->nv && nv_interior >= 0 && nv_interior <= nv_own, "bad partition sizes");
    h->nv_own = nv_own;
    h->nv_int = nv_interior;
    h->d_vecs.release();
    invalidate(h);
  });
}

int btfem_get_partition(btfem_t* h, int64_t* n_own, int64_t* n_interior, int64_t* halo_shift) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && h->nv_own >= 0, "partitioned, assembled handle required");
    if (n_own) *n_own = h->n_own;
    if (n_interior) *n_interior = h->n_int;
    if (halo_shift) *halo_shift = h->halo_shift;
  });
}

int btfem_dist_export(btfem_t* h, void* blob) {
  return guarded(h, [&] {
    BT_REQUIRE(h->assembled && blob, "call btfem_assemble first");
    bt_dist_export(h, blob);
  });
}

int btfem_dist_connect(btfem_t* h, int32_t rank, int32_t world, const void* blobs, int64_t nsend, const int32_t* src,
                       const int32_t* dst_rank, const int32_t* dst_slot, int64_t nsend_u, const int32_t* src_u,
                       const int32_t* dst_rank_u, const int32_t* dst_index_u, const int32_t* recv_from) {
  return guarded(h, [&] {
    BT_REQUIRE(blobs && recv_from && nsend >= 0 && (nsend == 0 || (src && dst_rank && dst_slot)), "null argument");
    BT_REQUIRE(nsend_u >= 0 && (nsend_u == 0 || (src_u && dst_rank_u && dst_index_u)), "null argument");
    bt_dist_connect(h, rank, world, blobs, nsend, src, dst_rank, dst_slot, nsend_u, src_u, dst_rank_u, dst_index_u,
                    recv_from);
  });
}

int btfem_dist_trace(btfem_t* h, int64_t max_entries) {
  return guarded(h, [&] {
    BT_REQUIRE(max_entries >= 0 && max_entries < (1LL << 28), "bad trace size");
    BT_REQUIRE(!h->dist_connected, "enable tracing before btfem_dist_connect");
    h->trace_cap = max_entries;
  });
}

int btfem_dist_get_trace(btfem_t* h, uint64_t* out, int64_t max_entries, int64_t* n_entries) {
  return guarded(h, [&] {
    BT_REQUIRE(out && n_entries && max_entries >= 0, "null argument");
    *n_entries = bt_dist_get_trace(h, out, max_entries);
  });
}

}  // extern "C"
