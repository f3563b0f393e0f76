/* btfem.h -- C-ABI of libbtfem.so, the B200 (sm_100a) Bloch-Torrey theta-scheme stepper.
 *
 * The reference (van-dang/DMRI-FEM-Cloud) has no FFI: its seam is the Python call
 *   MRI_simulation.solve(mydomain, mri_para, linsolver)        DmriFemLib.py:878-915
 *   PostProcessing(mydomain, mri_para, mri_simu, ...)          DmriFemLib.py:917-988
 * underneath which DOLFIN `assemble` and PETSc `KSPSolve` do all arithmetic.  This header
 * is what a ctypes stub in DmriFemLib.py would bind to replace that arithmetic; each entry
 * point cites the reference lines it stands for.  See INTEGRATION.md for the stub.
 *
 * Conventions: plain pointers and sizes only; every array is caller-owned, contiguous,
 * host memory, copied inside the call.  The handle is library-owned, bound to ONE GPU and
 * not thread-safe.  Every function returns 0 on success or a negative BTFEM_E* code;
 * btfem_last_error() then describes it.  No exception crosses the boundary.  There is no
 * CPU fallback: without a usable CUDA device btfem_create fails.
 *
 * Unknown numbering: one complex number per ACTIVE (vertex, compartment) pair, numbered
 * vertex-major ("dofs").  One compartment: dof == vertex.  Complex vectors are interleaved
 * (re, im) doubles.  All matrices are real, fp64, and share one CSR pattern with sorted
 * int32 columns.
 */
#ifndef BTFEM_H
#define BTFEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct btfem btfem_t;

enum {
  BTFEM_OK = 0,
  BTFEM_EINVAL = -1,     /* bad argument / call order */
  BTFEM_ECUDA = -2,      /* CUDA runtime error */
  BTFEM_ENOTCONV = -3,   /* Krylov: maximum iterations reached (KSP_DIVERGED_ITS) */
  BTFEM_EBREAKDOWN = -4, /* Krylov: rho == 0 or (v,r^) == 0 (KSP_DIVERGED_BREAKDOWN) */
  BTFEM_ENAN = -5,       /* Krylov: non-finite residual (KSP_DIVERGED_NANORINF) */
  BTFEM_EDTOL = -6,      /* Krylov: residual grew by dtol (KSP_DIVERGED_DTOL) */
  BTFEM_ENOMEM = -7,
  BTFEM_ECOMM = -8       /* row-partitioned solve: a peer did not answer in time (the solve is abandoned) */
};

/* which matrix btfem_get_values returns */
enum { BTFEM_MAT_M = 0, BTFEM_MAT_S = 1, BTFEM_MAT_R = 2, BTFEM_MAT_JX = 3, BTFEM_MAT_JY = 4,
       BTFEM_MAT_JZ = 5, BTFEM_MAT_I = 6, BTFEM_MAT_B = 7 };

enum { BTFEM_KSP_BICGSTAB = 0, BTFEM_KSP_GMRES = 1 };
/* BTFEM_PC_ILU: ILU(0) on the pattern of A (PETSc PCILU defaults: levels 0, natural ordering), the preconditioner of
 * KrylovSolver("gmres","ilu") (comri/one-comp/fenics-cpp/main.cpp:180-183) and PETSc's serial default; single
 * whole-mesh solves, zero initial guess.  The factorisation is renewed whenever theta*cA[step] changes. */
enum { BTFEM_PC_JACOBI = 0, BTFEM_PC_NONE = 1, BTFEM_PC_ILU = 2 };

/* ---- lifetime ----------------------------------------------------------------------- */
int btfem_create(int device, btfem_t** out);
void btfem_destroy(btfem_t* h);
const char* btfem_last_error(btfem_t* h);   /* valid until the next call on h */
int btfem_version(void);

/* ---- problem definition ---------------------------------------------------------------
 * Mesh + phase function: replaces Mesh/HDF5File.read + `phase` DG0 function
 * (GCloudDmriSolver.py:150-177; phase = marker % 2, DmriFemLib.py:764).
 * phase == NULL: one compartment (-M 0).  Otherwise phase[c] in {0,1} (-M 1). */
int btfem_set_mesh(btfem_t* h, int64_t nv, const double* xyz /*[nv*3]*/, int64_t nc,
                   const int32_t* tets /*[nc*4]*/, const int32_t* phase /*[nc] or NULL*/);

/* Triangle meshes: gdim-2 meshes (the 2-D disks of ArbitraryTimeSequence.ipynb / T2_Relaxation.ipynb; the caller
 * passes z = 0 and GdotX reduces to x*g0 + y*g1, DmriFemLib.py:34-36) and surfaces embedded in 3-D
 * (Manifolds.ipynb; tdim 2, gdim 3, DmriFemLib.py:591-592).  Same path afterwards: P1 element integrals on
 * triangles, interface and boundary facets are edges.  btfem_get_boundary_facets then reports the third vertex
 * of a facet as -1.  Whole-mesh handles only (no btfem_set_partition). */
int btfem_set_mesh_tri(btfem_t* h, int64_t nv, const double* xyz /*[nv*3]*/, int64_t nc,
                       const int32_t* tris /*[nc*3]*/, const int32_t* phase /*[nc] or NULL*/);

/* Curves in 3-D (neuron skeletons: Manifolds.ipynb runs `fru_M_100383_1D.xml`, tdim 1, gdim 3): P1 segments,
 * a vertex may join any number of them (branch points).  One compartment, Neumann ends. */
int btfem_set_mesh_seg(btfem_t* h, int64_t nv, const double* xyz /*[nv*3]*/, int64_t nc,
                       const int32_t* segs /*[nc*2]*/);

/* Strongly imposed pseudo-periodic BC (`IsDomainPeriodic = True` with a periodic direction: FuncF_sBC,
 * outer_interface, inner_interface, ThetaMethodF/L_sBC1c/2c and `constrained_domain = PeriodicBD`,
 * DmriFemLib.py:147-238, 327-375, 478-483).  vmaster[v] = the vertex on the min face that vertex v of a max face is
 * identified with (v itself otherwise; host search: periodic.vertex_map).  Slave vertices then carry their
 * master's dofs and btfem_solve steps the TRANSFORMED equation: pass cA[n] = q*F(t_n), cb[n] = q*F(t_{n-1}) (the
 * integrated profile, DmriFemLib.py:901-902) instead of q*f.  NULL: back to the default.  Call before
 * btfem_assemble.  BiCGStab, whole-mesh handles, tetrahedra or triangles.
 * (Compiled and restated by the oracle; GPU parity tests pending -- see csrc/strong.cu.) */
int btfem_set_periodic_map(btfem_t* h, const int32_t* vmaster /*[nv] or NULL*/);
/* parity hook: W[i,j] = int (g.Dg) phi_i phi_j and G = C - N of the transformed equation, CSR order */
int btfem_get_strong_operators(btfem_t* h, const double gdir[3], double* W /*[nnz] or NULL*/,
                               double* G /*[nnz] or NULL*/);

/* Replace only the phase function of the current mesh (NULL = one compartment). */
int btfem_set_phase(btfem_t* h, const int32_t* phase /*[nc] or NULL*/);
/* mesh.hmin()/hmax() as used by MyDomain (DmriFemLib.py:588-589): min / max over cells of the cell size,
 * cell size = longest edge (DOLFIN >= 2017 Cell::h, third party).  Computed on the GPU. */
int btfem_get_mesh_stats(btfem_t* h, double* hmin, double* hmax);
/* Bounding box of the vertices: GetGlobalDomainSize (DmriFemLib.py:560-581; MPI.min / MPI.max of the coordinates). */
int btfem_get_bbox(btfem_t* h, double lo[3], double hi[3]);

/* Diffusion: kind 0 = scalar D0 (`-K`, GCloudDmriSolver.py:212-215), 1 = per-cell scalar [nc],
 * 2 = per-cell full tensor [nc*9] row-major d00..d22 (ImposeDiffusionTensor, DmriFemLib.py:611-616). */
int btfem_set_diffusion(btfem_t* h, int kind, const double* D);
/* 1/T2: kind 0 = scalar, 1 = per-cell [nc] (DG0 `T2` function; FuncF_wBC, DmriFemLib.py:43-44). */
int btfem_set_relaxation(btfem_t* h, int kind, const double* inv_t2);
/* Membrane permeability kappa (`-p`; icondition_wBC, DmriFemLib.py:47-50): kind 0 = scalar,
 * 1 = table by cell-marker pair: kappa[(min(ma,mb))*nmark + max(ma,mb)], markers from
 * `marker` [nc] (variable permeability, MultilayeredDiskVariablePermeability.ipynb cell 10). */
int btfem_set_permeability(btfem_t* h, int kind, const double* kappa, int32_t nmark,
                           const int32_t* marker /*[nc] or NULL*/);
/* Weak pseudo-periodic BC (`-pdir`; DmriFemLib.py:58-77,79-93,256-324,599-610).
 * kappa_e: artificial permeability (reference: 3e-3/hmin); tol: face tolerance (1e-2*hmin);
 * bbox lo/hi as printed by MyDomain (DmriFemLib.py:593). */
int btfem_set_periodic(btfem_t* h, const int32_t pdir[3], double kappa_e, double tol,
                       const double lo[3], const double hi[3]);
/* Exterior facets that touch the periodic marker (the facets B is assembled on), as sorted vertex triples;
 * n_bfacet from btfem_get_sizes.  Lets the host build the periodic gather without its own facet search. */
int btfem_get_boundary_facets(btfem_t* h, int32_t* verts /*[n_bfacet*3]*/);
/* Gather operator of the weak pseudo-periodic BC (WeakPseudoPeriodic_*.eval, DmriFemLib.py:270-321): for
 * boundary dof `dof[b]`:  u_bc = exp(i*q*(g . dx[b])*F(t_p)) * sum_k w[b][k] * u[src[b][k]]   (src == -1: term is 0).
 * Built once per mesh by the host layer (periodic.build_gather); call after btfem_assemble.
 * Row-partitioned handles: dof / src are LOCAL dofs; src = -2-k means entry k of the periodic source buffer, which
 * the owning peer fills (btfem_dist_connect, `_u` list); call before btfem_dist_export. */
int btfem_set_periodic_gather(btfem_t* h, int64_t nb, const int32_t* dof /*[nb]*/, const int32_t* src /*[nb*3]*/,
                              const double* w /*[nb*3]*/, const double* dx /*[nb*3]*/);
/* Initial condition per vertex (Dirac_Delta interpolant, DmriFemLib.py:865-876); NULL = 1. */
int btfem_set_initial(btfem_t* h, const double* ic /*[nv] or NULL*/);

/* ---- assembly (once) -------------------------------------------------------------------
 * Replaces assemble(F), assemble(L), MassMatrix (DmriFemLib.py:240-254, 904-905) and the
 * comri pre-assembly of M, S, J[, I] (comri/one-comp/hpc-fenics-cpp/main.cpp:263-281):
 * builds the dof map, the CSR pattern (GPU sort/unique) and M,S,R,Jx,Jy,Jz,I,B. */
int btfem_assemble(btfem_t* h);

/* sizes after assemble: ndof, nnz, number of interface facets, number of boundary facets in B */
int btfem_get_sizes(btfem_t* h, int64_t* ndof, int64_t* nnz, int64_t* n_iface, int64_t* n_bfacet);
/* parity hooks */
int btfem_get_pattern(btfem_t* h, int32_t* rowptr /*[ndof+1]*/, int32_t* colidx /*[nnz]*/);
int btfem_get_dofmap(btfem_t* h, int32_t* dof_vertex /*[ndof]*/, int32_t* dof_comp /*[ndof]*/);
int btfem_get_values(btfem_t* h, int which, double* out /*[nnz]*/);
int btfem_get_lumped_mass(btfem_t* h, double* out /*[ndof]*/);   /* 1^T M: signal weights */

/* y = (P + i*c*Jg) x on host vectors, P = M/dt + theta*(S+R+I+B), Jg = g.J (g as given).
 * The product of `A = 1/k*M + assemble(F)` with a vector (DmriFemLib.py:904). */
int btfem_spmv(btfem_t* h, double dt, double theta, double c, const double gdir[3],
               const double* x /*[2*ndof]*/, double* y /*[2*ndof]*/);
/* Time `nrep` launches of the fused SpMV kernel on device-resident data (bench hook).
 * lanes: 0 = SELL-32 layout (default), or CSR with that many threads per row (4, 8, 16, 32).
 * flush_l2 != 0: a kernel reading a 512 MiB scratch buffer precedes every launch (evicts L2) and
 * each launch is timed on its own; otherwise the nrep launches are timed back to back.
 * Returns average ms per launch. */
int btfem_spmv_bench(btfem_t* h, double dt, double theta, double c, const double gdir[3],
                     int32_t lanes, int32_t nrep, int32_t flush_l2, double* ms_per_launch);
/* SpMV variant used by the solver: 0 = SELL-32 (default), else CSR with `lanes` threads per row */
int btfem_set_lanes(btfem_t* h, int32_t lanes);
/* Which fused-SpMV kernel a single whole-mesh solve on this (assembled) handle runs: 0 = CSR, `lanes` threads per row;
 * 1 = SELL-32, register-staged loads (k_spmv_sell); 2 = SELL-32 through per-warp TMA rings (k_spmv_stream:
 * cp.async.bulk + mbarrier).  Bench / profiling hook; the reference has no counterpart (PETSc MatMult). */
int btfem_get_spmv_kernel(btfem_t* h, int32_t* kind);
/* SM partition: the persistent time-loop kernel of this handle runs on `nblocks` SMs (one block each; 0 = all SMs of
 * the device).  Several handles with disjoint shares -- e.g. 16 x 9 SMs of a B200 -- solve CONCURRENTLY on one GPU,
 * each from its own host thread on its own stream: the serial loops over directions / b-values around `solve`
 * (ExplicitImplementation.ipynb cell 10) on meshes too small to fill the device.  Set before btfem_assemble (the
 * operator's warp-stream layout is built for this launch shape). */
int btfem_set_sm_partition(btfem_t* h, int32_t nblocks);
/* Parity hook: the ILU(0) factors of the last solve with BTFEM_PC_ILU, (re,im) per CSR nonzero (unit-lower L below the
 * diagonal, U on and above it -- one array, like PETSc's factored AIJ matrix).  out[2*nnz]. */
int btfem_get_ilu_factors(btfem_t* h, double* out);

/* ---- the theta loop ---------------------------------------------------------------------
 * MRI_simulation.solve (DmriFemLib.py:878-915): for n in 0..nsteps-1
 *   (P + i*theta*cA[n]*Jg) u^{n+1} = (Q - i*(1-theta)*cb[n]*Jg) u^n + (1-theta)*B*u_bc(u^n, Fb[n])
 * with cA[n] = q*f(t_n), cb[n] = q*f(t_{n-1}) (lagged, DmriFemLib.py:901-902,909), solved by
 * KrylovSolver("bicgstab","jacobi") semantics (GCloudDmriSolver.py:219-222; PETSc KSPBCGS,
 * left PCJACOBI, preconditioned-residual test max(rtol*||K^-1 b||, atol)).
 * All fields are 8 bytes wide so that the struct has no padding. */
typedef struct {
  int64_t nsteps;
  double dt;
  double theta;
  const double* cA;   /* [nsteps] */
  const double* cb;   /* [nsteps] */
  const double* Fb;   /* [nsteps] or NULL (only read with periodic BC) */
  double gdir[3];     /* unit gradient direction (set_gradient_dir normalises, DmriFemLib.py:819-821) */
  double q;           /* q-value, enters the periodic phase only */
  int64_t ksp;        /* BTFEM_KSP_* */
  int64_t pc;         /* BTFEM_PC_*  */
  double rtol;
  double atol;
  int64_t maxit;
  int64_t nonzero_guess;   /* parameters["krylov_solver"]["nonzero_initial_guess"] */
  int64_t restart;         /* GMRES restart (PETSc default 30) */
} btfem_solve_args;

typedef struct {
  double signal;          /* assemble(ur*dx) | assemble((phase*u1r+(1-phase)*u0r)*dx), DmriFemLib.py:931,971 */
  double signal_comp[2];  /* signal0, signal1 (DmriFemLib.py:928,930) */
  double voi;             /* assemble(Dirac_Delta*dx), DmriFemLib.py:924 */
  double voi_comp[2];     /* initial0, initial1 (DmriFemLib.py:927,929) */
  double whole_vol;       /* assemble(1*dx), DmriFemLib.py:923 */
  double loop_ms;         /* device time of the time loop (CUDA events) */
  double setup_ms;        /* device time of the per-solve operator combination */
  int64_t total_iters;    /* Krylov iterations summed over steps */
  int64_t max_iters;      /* largest per-step count */
  int64_t n_spmv;         /* fused-SpMV launches that did work */
  int64_t n_kernels;      /* all kernel launches inside the loop */
  int64_t last_reason;    /* >0 converged (2 = rtol, 3 = atol) */
} btfem_solve_out;

int btfem_solve(btfem_t* h, const btfem_solve_args* args, btfem_solve_out* out,
                int32_t* iters_per_step /*[nsteps] or NULL*/);
/* `members` independent solves on the same mesh in lock step (HARDI sweeps: the reference loops serially over
 * directions and b-values, ExplicitImplementation.ipynb cell 10): args[m] differ in gdir and cA/cb (q); nsteps, dt,
 * theta and the Krylov settings are taken from args[0].  BiCGStab, no periodic BC.  Up to 32 members on a whole-mesh
 * handle run their whole time loop as ONE cooperative kernel launch (members of one direction share an operator copy;
 * a member that has converged in a time step drops out until the next one; a failing member ends the batch and the
 * call returns its error); larger batches advance through one kernel chain per iteration that covers all members.
 * Results are reproducible run to run; which batch a member travels in changes its dot-product grouping (agreement
 * with the one-at-a-time solve to solver tolerance, 1e-9 relative in the tests). */
int btfem_solve_batch(btfem_t* h, int32_t members, const btfem_solve_args* args /*[members]*/,
                      btfem_solve_out* out /*[members]*/);
/* solution after the last solve, dof numbering, interleaved (re,im) */
int btfem_get_solution(btfem_t* h, double* u /*[2*ndof]*/);

/* ---- one mesh row-partitioned over several GPUs ---------------------------------------------
 * Replaces the MPI domain decomposition that DOLFIN/PETSc do underneath `mpirun -n N python3 GCloudDmriSolver.py`
 * (README.md:89-94; MatMult ghost scatter + MPI_Allreduce inside KSPSolve, third party).  One process (or
 * thread) per GPU; every rank owns a contiguous block of vertices of a locality-ordered numbering and builds
 * a handle on its LOCAL sub-mesh: all cells touching an owned vertex, vertices numbered
 *   [ owned, not needed by peers | owned, needed by peers | halo (owned by peers) ].
 * Rows of owned dofs are complete; rows of halo dofs are never used.  During the solve there is no exchange
 * step: the kernel that produces a Krylov vector stores the entries peers need straight into the peers' halo
 * slots (NVLink peer stores), the fused SpMV starts on rows that need no halo at once and waits for the peers'
 * arrival flags only where it reaches the first row that does; the dot products are all-reduced by the last
 * thread block of the producing kernel through peer memory (fixed rank order -> every rank holds bit-identical
 * scalars, so all ranks take identical control decisions).  BiCGStab only; Neumann or weak pseudo-periodic BC.
 * Call order: set_mesh / coefficients -> btfem_set_partition -> btfem_assemble -> btfem_dist_export ->
 * (host layer all-gathers the blobs and the halo requests) -> btfem_dist_connect -> host barrier -> btfem_solve.
 * btfem_solve then returns the GLOBAL signal on every rank; voi / whole_vol are sums over OWNED dofs (the
 * host layer adds them up). */
#define BTFEM_DIST_BLOB_BYTES 192
int btfem_set_partition(btfem_t* h, int64_t nv_own, int64_t nv_interior /* <= nv_own */);
/* after assemble: owned dofs, owned dofs not needed by peers, and the element shift of halo dofs in vectors */
int btfem_get_partition(btfem_t* h, int64_t* n_own, int64_t* n_interior, int64_t* halo_shift);
int btfem_dist_export(btfem_t* h, void* blob /*[BTFEM_DIST_BLOB_BYTES]*/);
/* blobs: every rank's blob, rank-major.  Send list: owned dof src[e] goes to vector element dst_slot[e]
 * (= peer-local dof + the peer's halo_shift) of rank dst_rank[e] whenever u, p or s is exchanged.
 * The `_u` list travels with u only (once per time step): owned dof src_u[e] goes to entry dst_index_u[e] of
 * rank dst_rank_u[e]'s periodic source buffer -- the weak pseudo-periodic gather reads u at mirrored points,
 * which a peer may own (btfem_set_periodic_gather: src <= -2).  recv_from[r] != 0 if rank r sends to us. */
int btfem_dist_connect(btfem_t* h, int32_t rank, int32_t world, const void* blobs, int64_t nsend,
                       const int32_t* src, const int32_t* dst_rank, const int32_t* dst_slot, int64_t nsend_u,
                       const int32_t* src_u, const int32_t* dst_rank_u, const int32_t* dst_index_u,
                       const int32_t* recv_from /*[world]*/);

/* Optional device-side timeline of a row-partitioned solve (measurement aid): call btfem_dist_trace before
 * btfem_dist_connect; every kernel that closes a collective (halo publish in the update kernels, all-reduce in the
 * SpMV / update_xr / signal kernels) then logs 8 words {first block start, local work done, collective done, longest
 * halo wait, kind, 0, 0, 0} (GPU globaltimer ns; kind: 1 RHS, 2 residual, 3 SpMV v=Ap, 4 SpMV t=As, 5 update p,
 * 6 update s, 7 update x/r, 0 other), for the first max_entries such kernels. */
int btfem_dist_trace(btfem_t* h, int64_t max_entries);
int btfem_dist_get_trace(btfem_t* h, uint64_t* out /*[max_entries*8]*/, int64_t max_entries, int64_t* n_entries);

#ifdef __cplusplus
}
#endif
#endif /* BTFEM_H */
