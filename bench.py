#!/usr/bin/env python3
"""Benchmark of the Bloch-Torrey theta-scheme path (BASELINE.json metric: DOF-steps/s).

Workload (config.workload): BASELINE.json configs[1], the two-compartment permeable PGSE solve
`-M 1 -b 1000 -p 1e-5 -k 200 -gdir 0 1 0` (delta/Delta = 10600/43100 -> 270 theta steps,
D = 3e-3, BiCGStab+Jacobi rtol 1e-9 atol 1e-10, GCloudDmriSolver.py:52-55,219-222) on the
synthetic cell-in-box mesh (sphere R=5 in a [-10,10]^3 Kuhn box) sized to ~1 M real DOFs.

One bench "step" = one complete solve (IC -> 270 theta steps -> signal).  DOFs = 2 x active
complex unknowns (the reference would count W.dim() = 4 x N_vert = 2x more for two compartments).

  value    throughput of btfem_solve with mesh + operators resident in HBM (wall clock around
           K solves, each ending in a stream sync; device event time reported beside it)
  e2e      same metric through the reference-facing API (MyDomain / MRI_simulation.solve /
           PostProcessing) from HOST numpy buffers: mesh upload, dof map, pattern, assembly,
           solve, signal read-back all inside the timed region
  roofline fused complex SpMV kernel: algorithmic bytes 20*nnz + 36*n per launch over the
           CUDA-event time of one launch with L2 flushed before it
  cpu_baseline  the oracle's C/OpenMP restatement on a bounded sample of the same workload
  loop_roofline the whole Krylov loop (algorithmic bytes of all iterations over the device loop time)
  hardi    BASELINE.json's third figure, dMRI signals/s: the 64 x 4 HARDI sweep of configs[4] sharded over the
           N ranks (skip with --no-hardi); a failure there is reported in the key and never blocks the line
  partitioned  N > 1 only: ONE mesh row-partitioned over the N GPUs (configs[3]: ECS of 226 cylinders, weak
           pseudo-periodic BC, ~1 M and ~4 M DOFs; the reference's `mpirun -n N` mode, README.md:86-94) -- DOF-steps/s,
           us per BiCGStab iteration, speed-up and relative signal difference against the same solve on one GPU
  signal_check  N = 1 only: the oracle's C/OpenMP restatement runs ALL theta steps of the same problem at the same
           Krylov tolerances on the host cores; the relative difference of the two final signals is printed
           (north star: <= 1e-8) and the same run is the `cpu_baseline`

N > 1 (torchrun): independent gradient directions shard one per rank, no data-path collective
(weak scaling); torch.distributed is used for the barrier and the max-over-ranks only.
`--impl reference` times the CPU restatement (the reference itself cannot run here: no
DOLFIN/PETSc/MPI in the image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "Bloch-Torrey DOF-steps/s"
UNIT = "DOF-steps/s"
# DRAM bytes of one k_spmv_sell<MODE_V> launch on the default workload, from the committed
# `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_ncu_spmv_sell_full.txt)
NCU_TRAFFIC = {"n_box": 78,
               "sell": {"bytes": 167.227136e6 + 3.839232e6, "file": "profiles/r1_ncu_spmv_sell_full.txt (ncu, round 1)"},
               # one launch = one whole solve of this workload (270 steps, 11 335 iterations when captured; the final layout:
               # 16-bit column offsets, 18 B per nonzero)
               "persistent": {"bytes": 3.117732e12 + 95.638328e9,
                              "file": "profiles/r2m_ncu_persistent_and_stream_full.txt (ncu, round 2, last capture; 0.63 x the "
                                      "algorithmic bytes: the Krylov vectors stay in L2, HBM carries the operator stream)"}}


def workload(n_box):
    entry.load_package()
    from dmri_fem_cloud_b200 import meshes
    xyz, tets, phase = meshes.box_with_sphere(10.0, n_box, 5.0)
    return xyz, tets, phase


def sequence(delta=10600.0, Delta=43100.0, k=200.0, b=1000.0):
    """PGSE scalars exactly as GCloudDmriSolver.py:184-193 + DmriFemLib.py:826-858 produce them."""
    entry.load_package()
    from dmri_fem_cloud_b200 import dmrifemlib as dl
    import sympy as sp
    mp = dl.MRI_parameters()
    mp.bvalue = b
    mp.delta, mp.Delta = delta, Delta
    mp.T = Delta + delta
    mp.fs_sym = sp.Piecewise((1., mp.s < delta), (0., mp.s < Delta), (-1., mp.s < mp.T), (0., True))
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = k
    ts = sim.time_grid(mp)
    f, _ = mp.profiles_on_grid(ts)
    fp = np.concatenate([[f[0]], f[:-1]])
    return mp, ts, f, fp


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_operators(xyz, tets, phase):
    """The oracle's assembled operators of the bench workload (numpy; built once per process)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bt_oracle as orc
    return orc.assemble(xyz, tets, phase, D=3e-3, invT2=1e-16, kappa=1e-5)


def cpu_restatement(xyz, tets, phase, mp, f, fp, k, sample_steps, mode=0, ops=None):
    """Time the oracle's C/OpenMP time loop on the first `sample_steps` theta steps."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bt_cpu
    bt_cpu.set_threads(len(os.sched_getaffinity(0)))      # all host cores, whatever OMP_NUM_THREADS says
    if ops is None:
        ops = cpu_operators(xyz, tets, phase)
    q = mp.qvalue
    t0 = time.perf_counter()
    if mode == 2:      # DmriFemLib.solve work pattern: A and b re-assembled from the elements every step
        u, iters = bt_cpu.theta_loop_reassemble(ops, xyz, tets, [0, 1, 0], k, 0.5, q * f[:sample_steps],
                                                q * fp[:sample_steps], 3e-3, 1e-16, 1e-5)
    else:
        u, iters = bt_cpu.theta_loop(ops, [0, 1, 0], k, 0.5, q * f[:sample_steps], q * fp[:sample_steps], mode=mode)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "ndof_real": 2 * ops.ndof, "steps": sample_steps, "iters": int(iters.sum()),
            "cores": bt_cpu.num_threads(), "signal": float(ops.lumped @ u.real), "voi": float(ops.lumped.sum()),
            "nnz": int(ops.nnz)}


def make_config(n_box, world, ndof_real, n_vertices, n_tets, nnz, nsteps):
    """The workload description BOTH arms print (the driver compares the two dicts)."""
    return {"workload": "configs[1] two-compartment permeable PGSE (-M 1 -b 1000 -p 1e-5 -k 200 -gdir 0 1 0), "
                        "cell-in-box n_box=%d" % n_box,
            "ndof_real": int(ndof_real), "n_vertices": int(n_vertices), "n_tets": int(n_tets), "nnz": int(nnz),
            "theta_steps_per_solve": int(nsteps), "krylov": "bicgstab+jacobi rtol 1e-9 atol 1e-10",
            "l2_policy": "working set (operator %.0f MB + 8 vectors %.0f MB) larger than L2; the roofline launch is "
                         "timed after reading a 512 MiB scratch buffer (L2 flush)" % (20.0 * nnz / 1e6,
                                                                                    8 * 8.0 * ndof_real / 1e6),
            "parallelism": "sweep-sharded x%d (one gradient direction per GPU)" % world}


def loop_roofline(nnz, n, iters, nsteps, loop_ms, peak_gbs):
    spmv = 20.0 * nnz + 36.0 * n
    per_iter = 2.0 * spmv + 224.0 * n
    total = iters * per_iter + nsteps * spmv
    achieved = total / (loop_ms * 1e-3) / 1e9
    return {"algorithmic_bytes_per_iteration": per_iter, "algorithmic_bytes_per_solve": total,
            "us_per_iteration": 1e3 * loop_ms / max(iters, 1), "achieved": achieved, "unit": "GB/s",
            "frac": achieved / peak_gbs,
            "note": "Krylov vectors (7 x 16 n bytes) stay in L2 through a persisting window, so part of this traffic "
                    "never reaches HBM"}


def hardi_sweep(local_rank, rank, world, batch=16, h=0.7, ndir=64):
    """BASELINE.json's third figure: dMRI signals/s of the HARDI sweep (configs[4]: 64 directions x 4 b-values on a
    46 k-vertex neuron-like mesh, PGSE 10600/43100, dt 200), sharded over the ranks with no collective in the loop
    (sweep.shard_balanced) and `batch` members per kernel launch.  Returns this rank's wall time for its share (the
    caller takes the max over ranks) and the number of signals of the whole job."""
    entry.load_package()
    from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes, sweep
    xyz, tets = meshes.neuron_like(h=h)
    xyz, tets = meshes.rcm_order(*meshes.shuffle_vertices(xyz, tets, 0))
    mp, _, _, _ = sequence(k=200.0)
    sim = dl.MRI_simulation()
    sim.k = 200.0
    dirs = meshes.fibonacci_hemisphere(ndir)
    bvals = [1000.0, 2000.0, 3000.0, 4000.0]
    par = dict(rtol=1e-9, atol=1e-10, maxit=100000)
    with btfem.BTFem(local_rank) as fem:
        fem.set_mesh(xyz, tets)
        fem.set_diffusion(3e-3)
        fem.set_relaxation(1e-16)
        fem.assemble()
        sweep.run_sweep(fem, mp, sim, dirs[:2], bvals[:2], par, batch=4)          # warm-up
        t0 = time.perf_counter()
        mine, sig = sweep.run_sweep(fem, mp, sim, dirs, bvals, par, rank=rank, world=world, batch=batch)
        dt = time.perf_counter() - t0                                           # solve_batch returns after a stream sync
    return dt, len(dirs) * len(bvals), int(len(xyz)), float(np.sum(sig))


def partitioned_solve(dist, local_rank, rank, world, ecs):
    """BASELINE.json configs[3]: the extracellular space of 226 cylinders (two compartments, D = 2e-3, kappa = 1e-5,
    delta/Delta = 10000/13000, dt 200, g = (1,1,0)/sqrt 2, b = 1000, weak pseudo-periodic BC in x and y) as ONE mesh
    row-partitioned over the `world` GPUs (partition.DistBTFem: halo entries and dot products travel through peer
    memory inside libbtfem's kernels), against the same solve on one GPU (rank 0).  All ranks call this; rank 0 gets
    the figures.  Loop time is device time (CUDA events inside libbtfem), max over ranks."""
    import datetime
    entry.load_package()
    from dmri_fem_cloud_b200 import btfem, meshes, partition, periodic
    group = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=600))
    comm = partition.TorchComm(dist, group=group)
    xyz, tets, phase = meshes.ecs_slab(ecs, ecs, 2)
    xyz, tets = meshes.coordinate_order(xyz, tets, axes=(1, 2, 0))       # vertex blocks = slabs normal to y
    mp, ts, f, fp = sequence(delta=10000.0, Delta=13000.0, k=200.0, b=1000.0)
    _, Fb = mp.profiles_on_grid(np.concatenate([[0.0], ts[:-1]]))
    g = np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
    pdir, k, q = [1, 1, 0], 200.0, mp.qvalue
    kw = dict(rtol=1e-9, atol=1e-10, maxit=100000, q=q, Fb=Fb)
    lo, hi = xyz.min(axis=0), xyz.max(axis=0)
    d = partition.DistBTFem(xyz, tets, comm, device=local_rank, phase=phase)
    try:
        d.set_diffusion(2e-3)
        d.set_relaxation(1e-16)
        d.set_permeability(1e-5)
        hmin, _ = d.mesh_stats()
        d.set_periodic(pdir, 3e-3 / hmin, 1e-2 * hmin, lo, hi)            # kappa_e, tol as MyDomain sets them
        d.assemble()
        d.solve(k, 0.5, q * f, q * fp, g, **kw)                           # warm-up
        comm.barrier()
        res = d.solve(k, 0.5, q * f, q * fp, g, **kw)
        loop_s = float(comm.max([res["loop_ms"] + res["setup_ms"]])[0]) * 1e-3
        ndof_real = 2 * d.ndof_global
    finally:
        d.close()
    out = None
    if rank == 0:
        with btfem.BTFem(local_rank) as fem:
            fem.set_mesh(xyz, tets, phase)
            fem.set_diffusion(2e-3)
            fem.set_relaxation(1e-16)
            fem.set_permeability(1e-5)
            fem.set_periodic(pdir, 3e-3 / hmin, 1e-2 * hmin, lo, hi)
            fem.assemble()
            dv, dc = fem.dofmap()
            fem.set_periodic_gather(*periodic.build_gather(xyz, tets, phase, pdir, lo, hi, dv, dc,
                                                           bfacets=fem.boundary_facets()))
            fem.solve(k, 0.5, q * f, q * fp, g, **kw)
            ref = fem.solve(k, 0.5, q * f, q * fp, g, **kw)
        single_s = (ref["loop_ms"] + ref["setup_ms"]) * 1e-3
        s_part, s_one = res["signal"] / res["voi"], ref["signal"] / ref["voi"]
        out = {"workload": "configs[3] ECS slab %dx%dx2, 226 cylinders, weak periodic x,y, row-partitioned over %d GPUs"
                           % (ecs, ecs, world),
               "ndof_real": int(ndof_real), "theta_steps": len(ts), "iters": int(res["total_iters"]),
               "loop_s": loop_s, "value": ndof_real * len(ts) / loop_s, "unit": UNIT,
               "us_per_iteration": 1e6 * loop_s / max(1, int(res["total_iters"])),
               "single_gpu_loop_s": single_s, "single_gpu_iters": int(ref["total_iters"]),
               "speedup_vs_single_gpu": single_s / loop_s,
               "rel_signal_err_vs_single_gpu": abs(s_part - s_one) / abs(s_one)}
    comm.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="btfem", choices=["btfem", "reference"])
    ap.add_argument("--n-box", type=int, default=78, help="cubes per edge of the cell-in-box mesh (78 -> ~1 M DOFs)")
    ap.add_argument("--cpu-sample-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-hardi", action="store_true", help="skip the HARDI signals/s figure")
    ap.add_argument("--no-full", action="store_true", help="skip the full-length CPU run (signal check / reference arm)")
    ap.add_argument("--no-partitioned", action="store_true", help="N > 1: skip the row-partitioned single-mesh figure")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    k = 200.0

    if args.impl == "reference":
        if rank != 0:
            return 0
        entry.build_oracle()
        xyz, tets, phase = workload(args.n_box)
        mp, ts, f, fp = sequence(k=k)
        sample = args.cpu_sample_steps
        ops = cpu_operators(xyz, tets, phase)          # numpy assembly of the pattern: once, outside the timed region
        res = None
        for _ in range(max(1, args.warmup > 0)):       # one untimed pass pages everything in
            res = cpu_restatement(xyz, tets, phase, mp, f, fp, k, 1, mode=2, ops=ops)
        tsum, steps = 0.0, 0
        for _ in range(args.steps):
            res = cpu_restatement(xyz, tets, phase, mp, f, fp, k, sample, mode=2, ops=ops)
            tsum += res["seconds"]
            steps += sample
        v = res["ndof_real"] * steps / tsum
        full = None
        if not args.no_full:                           # one full-length solve beside the sampled steps
            r = cpu_restatement(xyz, tets, phase, mp, f, fp, k, len(ts), mode=2, ops=ops)
            full = {"theta_steps": len(ts), "seconds": r["seconds"], "value": r["ndof_real"] * len(ts) / r["seconds"],
                    "iters": r["iters"], "normalized_signal": r["signal"] / r["voi"]}
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tsum / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": make_config(args.n_box, args.gpus, res["ndof_real"], len(xyz), len(tets), res["nnz"], len(ts)),
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": res["cores"], "kind": "port",
                                 "sample": "first %d of %d theta steps per bench step; C/OpenMP restatement of "
                                           "DmriFemLib.solve: A and b re-assembled from element integrals every step "
                                           "(closed forms, cheaper than the reference's FFC kernels) + Jacobi-BiCGStab "
                                           "at the CLI tolerances; FEniCS/PETSc itself is not installable here" % (
                                               sample, len(ts)),
                                 "full_length_run": full},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ GPU arm
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    entry.load_package()
    from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes

    xyz, tets, phase = workload(args.n_box)
    mp, ts, f, fp = sequence(k=k)
    q = mp.qvalue
    # sweep sharding: rank r takes direction r of a fixed set (rank 0: the CLI's -gdir 0 1 0)
    dirs = np.vstack([[0.0, 1.0, 0.0], meshes.fibonacci_hemisphere(max(world, 2))])
    g = dirs[rank % len(dirs)]
    g = g / np.linalg.norm(g)

    fem = btfem.BTFem(local_rank)
    if args.lanes:
        fem.set_lanes(args.lanes)
    fem.set_mesh(xyz, tets, phase)
    fem.set_diffusion(3e-3)
    fem.set_relaxation(1e-16)
    fem.set_permeability(1e-5)
    t0 = time.perf_counter()
    fem.assemble()
    assemble_s = time.perf_counter() - t0
    ndof_real = 2 * fem.ndof
    nsteps = len(ts)

    def one_solve():
        return fem.solve(k, 0.5, q * f, q * fp, g, rtol=1e-9, atol=1e-10, maxit=100000)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = one_solve()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    kernels = 0
    for _ in range(args.steps):
        res = one_solve()                       # returns after a stream sync
        dev_ms += res["loop_ms"] + res["setup_ms"]
        kernels += res["n_kernels"] + 3
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e through the reference-facing API, host buffers -> signal
    def e2e_once():
        t_begin = time.perf_counter()
        mesh = dl.Mesh(xyz, tets)
        mesh.device = local_rank
        md = dl.MyDomain(mesh, mp)
        md.phase = phase
        md.IsDomainMultiple = True
        md.kappa = 1e-5
        md.Apply()
        md.D0 = 3e-3
        md.D = md.D0
        ls = dl.KrylovSolver("bicgstab", "jacobi")
        ls.parameters["relative_tolerance"] = 1e-9
        ls.parameters["absolute_tolerance"] = 1e-10
        ls.parameters["maximum_iterations"] = 100000
        sim = dl.MRI_simulation()
        sim.k = k
        sim.verbose = False
        t_solve = time.perf_counter()
        sim.solve(md, mp, ls)
        s = sim.stats["signal"] / sim.stats["voi"]
        # where the step went: host + device set-up before solve(), the solve() call, and inside it the device loop
        e2e_stats.append({"before_solve_s": round(t_solve - t_begin, 4), "solve_call_s": round(time.perf_counter() - t_solve, 4),
                          "device_loop_s": round(1e-3 * sim.stats.get("loop_ms", 0.0), 4),
                          "device_setup_s": round(1e-3 * sim.stats.get("setup_ms", 0.0), 4)})
        keep.append(sim.fem)        # teardown is not part of the reference's timed region either (measured: closing inside
        return s                    # the region costs 0.14 s per solve); the caller closes the handle between two steps

    import contextlib
    import io
    e2e_steps = max(1, min(args.steps, 2))
    keep = []
    e2e_stats, e2e_times = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        mp.set_gradient_dir(None, *g)
        e2e_once()
        keep.pop().close()
        # every step is timed on its own (host buffers in -> normalized signal out, which ends with the device -> host
        # read of the signal); the handle is closed BETWEEN two steps, untimed, so that the next step's device arrays come
        # from the stream-ordered pool instead of fresh driver allocations (one 1.3 s outlier was seen with two live handles)
        e2e_sampler = ClockSampler(local_rank)      # the e2e steps get their own clock / throttle record
        if rank == 0:
            e2e_sampler.start()
        e2e_elapsed = 0.0
        for _ in range(e2e_steps):
            barrier()
            t1 = time.perf_counter()
            e2e_sig = e2e_once()
            e2e_times.append(time.perf_counter() - t1)
            e2e_elapsed += e2e_times[-1]
            keep.pop().close()
        barrier()
        e2e_clocks = e2e_sampler.stop() if rank == 0 else None
    if os.environ.get("BENCH_DEBUG") and rank == 0:
        sys.stderr.write("[bench] e2e steps (s): %s | incl. warm-up: %s\n" % (["%.3f" % t for t in e2e_times], e2e_stats))

    # ---- HARDI sweep (signals/s).  No collective inside: a rank that fails reports an infinite time, so the
    # max over ranks below cannot hang on it.
    hardi_dt, hardi_info = float("inf"), None
    if not args.no_hardi:
        try:
            barrier()
            with contextlib.redirect_stdout(io.StringIO()):
                hardi_dt, n_sig, n_vert, checksum = hardi_sweep(local_rank, rank, world)
            hardi_info = {"n_signals": n_sig, "n_vertices": n_vert, "batch": 16, "local_checksum": checksum}
        except Exception as exc:      # the headline line must still be printed
            hardi_dt, hardi_info = float("inf"), {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- one mesh row-partitioned over the N GPUs (configs[3]); never blocks the headline line
    part_info = None
    if dist is not None and not args.no_partitioned:
        part_info = {}
        for ecs in (400, 800):
            try:
                barrier()
                with contextlib.redirect_stdout(io.StringIO()):
                    r = partitioned_solve(dist, local_rank, rank, world, ecs)
                part_info["ecs%d" % ecs] = r
            except Exception as exc:
                part_info["ecs%d" % ecs] = {"error": "%s: %s" % (type(exc).__name__, exc)}
                break

    tmax, e2e_max, hardi_max = elapsed, e2e_elapsed, hardi_dt
    if dist is not None:
        import torch
        tt = torch.tensor([elapsed, e2e_elapsed, hardi_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tmax, e2e_max, hardi_max = float(tt[0]), float(tt[1]), float(tt[2])

    if rank == 0:
        value = world * ndof_real * nsteps * args.steps / tmax
        e2e_value = world * ndof_real * nsteps * e2e_steps / e2e_max
        # roofline of the dominant kernel, measured live (CUDA events inside libbtfem on its stream)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        lanes = args.lanes
        alg_bytes = 20.0 * fem.nnz + 36.0 * fem.ndof
        ms_cold = fem.spmv_bench(k, 0.5, q, g, lanes=lanes, nrep=20, flush_l2=True)
        ms_warm = fem.spmv_bench(k, 0.5, q, g, lanes=lanes, nrep=50, flush_l2=False)
        achieved = alg_bytes / (ms_cold * 1e-3) / 1e9
        stream = lanes == 0 and fem.stream_kernel
        spmv_kernel = ("k_spmv_stream<MODE_V|MODE_T> (SELL-32 through per-warp TMA rings)" if stream else
                       "k_spmv_sell<MODE_V|MODE_T>" if lanes == 0 else "k_spmv<%d,MODE_V|MODE_T>" % lanes)
        # DRAM bytes per launch: from the committed `ncu --set full` capture of this kernel on this workload
        # (profiles/, dram__bytes_read.sum + dram__bytes_write.sum); null when no capture matches the configuration
        traffic, traffic_src = None, None
        tkey = "stream" if stream else ("sell" if lanes == 0 else None)
        if args.n_box == NCU_TRAFFIC["n_box"] and tkey in NCU_TRAFFIC:
            traffic, traffic_src = NCU_TRAFFIC[tkey]["bytes"], NCU_TRAFFIC[tkey]["file"]
        spmv_share = res["n_spmv"] * ms_warm / max(res["loop_ms"], 1e-9)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * tmax / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": make_config(args.n_box, world, ndof_real, len(xyz), len(tets), fem.nnz, nsteps),
                "solve": {"n_interface_facets": fem.n_iface, "iters_per_solve": res["total_iters"],
                          "spmv_lanes_per_row": lanes, "spmv_kernel": spmv_kernel},
                "device_ms_per_step": dev_ms / args.steps, "assemble_s": assemble_s,
                "normalized_signal": res["signal"] / res["voi"],
                "gpu_launches": int(kernels),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(fem.h2d_bytes + 3 * 8 * nsteps),
                        "d2h_bytes_per_step": int(8 * 8 + 4 * 8), "seconds_per_solve": e2e_max / e2e_steps,
                        "rank0_steps": [dict(st, total_s=round(t, 4)) for st, t in zip(e2e_stats[1:], e2e_times)],
                        "clocks": e2e_clocks,
                        "normalized_signal": e2e_sig,
                        "api": "dmrifemlib.MyDomain/MRI_simulation.solve (host numpy mesh -> signal)"},
                "roofline": None}
        # ---- roofline of the dominant kernel.  With the persistent path ONE launch (k_bicgstab_persistent) is the whole
        # theta loop: its algorithmic bytes are SURVEY 8(d)'s per-unit figures x the units it processed (iterations x
        # [2 SpMV + 224 n of vector passes] + time steps x 1 SpMV for the right-hand side), its duration the CUDA-event
        # time libbtfem takes around it on its stream.  The fused SpMV alone (k_spmv_stream, the pass body of the
        # persistent kernel launched on its own, L2 flushed) is reported beside it as `spmv`.
        spmv_roof = {"kernel": spmv_kernel + " fused complex SpMV", "achieved": achieved, "unit": "GB/s",
                     "frac": achieved / peak, "algorithmic_bytes_per_launch": alg_bytes,
                     "ms_per_launch_l2_flushed": ms_cold, "ms_per_launch_back_to_back": ms_warm,
                     "achieved_back_to_back": alg_bytes / (ms_warm * 1e-3) / 1e9, "traffic": traffic,
                     "traffic_source": traffic_src}
        persistent = stream and res["n_kernels"] <= 4 + 3 * nsteps
        if persistent:
            lr = loop_roofline(fem.nnz, fem.ndof, res["total_iters"], nsteps, res["loop_ms"], peak)
            ptraf = NCU_TRAFFIC.get("persistent") if args.n_box == NCU_TRAFFIC["n_box"] else None
            line["roofline"] = {"bound": "hbm",
                                "kernel": "k_bicgstab_persistent (one cooperative launch = the whole theta loop of a solve: "
                                          "right-hand sides + Jacobi-BiCGStab iterations; SpMV passes on per-warp TMA rings)",
                                "achieved": lr["achieved"], "peak": peak, "unit": "GB/s", "frac": lr["frac"],
                                "peak_source": peak_src,
                                "traffic": ptraf["bytes"] if ptraf else None,
                                "traffic_source": ptraf["file"] if ptraf else None,
                                "algorithmic_bytes_per_launch": lr["algorithmic_bytes_per_solve"],
                                "ms_per_launch": res["loop_ms"], "us_per_iteration": lr["us_per_iteration"],
                                "l2_note": "one launch streams the 148 MB operator ~%d times: no L2 flush needed; the "
                                           "Krylov vectors (7 x 16 n bytes) are kept in L2 by evict-first operator loads, "
                                           "so part of the algorithmic vector traffic never reaches HBM" % (
                                               2 * res["total_iters"] + nsteps),
                                "spmv": spmv_roof}
        else:
            line["roofline"] = dict({"bound": "hbm", "peak": peak, "peak_source": peak_src,
                                     "share_of_loop": spmv_share}, **spmv_roof)
        # the whole time loop against the same roofline: SURVEY 8(d) algorithmic bytes of a Jacobi-BiCGStab iteration
        # (2 SpMVs + 224 n of vector passes) and of the per-step right-hand side (1 SpMV), over the device loop time
        if hardi_info is not None:
            if np.isfinite(hardi_max) and "error" not in hardi_info:
                hardi_info.update({"value": hardi_info["n_signals"] / hardi_max, "unit": "signals/s",
                                   "seconds": hardi_max,
                                   "workload": "configs[4] HARDI 64 directions x 4 b-values, neuron-like mesh, sharded over "
                                               "%d GPU(s), no collective in the loop" % world})
            elif "error" not in hardi_info:
                hardi_info["error"] = "a rank failed"
            line["hardi"] = hardi_info
        line["loop_roofline"] = loop_roofline(fem.nnz, fem.ndof, res["total_iters"], nsteps, res["loop_ms"], peak)
        if not args.no_cpu and world == 1:
            entry.build_oracle()
            ops = cpu_operators(xyz, tets, phase)
            cpu_steps = args.cpu_sample_steps if args.no_full else nsteps
            cpu = cpu_restatement(xyz, tets, phase, mp, f, fp, k, cpu_steps, mode=0, ops=ops)
            line["cpu_baseline"] = {"value": cpu["ndof_real"] * cpu["steps"] / cpu["seconds"], "unit": UNIT,
                                    "cores": cpu["cores"], "kind": "port",
                                    "sample": "%s %d of %d theta steps, oracle C/OpenMP restatement with "
                                              "pre-combined operators (%.1f s)" % (
                                                  "all" if cpu_steps == nsteps else "first", cpu["steps"], nsteps,
                                                  cpu["seconds"])}
            if cpu_steps == nsteps:      # same problem, same Krylov tolerances, all theta steps: the two final signals
                s_cpu, s_gpu = cpu["signal"] / cpu["voi"], res["signal"] / res["voi"]
                line["signal_check"] = {"signal_rel_err_vs_cpu": abs(s_gpu - s_cpu) / abs(s_cpu),
                                        "normalized_signal_gpu": s_gpu, "normalized_signal_cpu": s_cpu,
                                        "krylov_rtol": 1e-9, "krylov_atol": 1e-10, "theta_steps": nsteps,
                                        "iters_gpu": int(res["total_iters"]), "iters_cpu": cpu["iters"],
                                        "target": 1e-8}
        else:
            line["cpu_baseline"] = None
        if part_info is not None:
            line["partitioned"] = part_info
        print(json.dumps(line))
    fem.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
