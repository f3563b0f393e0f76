"""One mesh row-partitioned over the GPUs of a node, one process per GPU (run under torchrun):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/dist_solve.py --nbox 78 [--check] [--reps 2] [--ecs NX]

Workload: bench.py's configs[1] cell-in-box PGSE solve (or, with --ecs, the two-compartment extracellular-space
slab of configs[3] without its periodic BC).  The host side only needs python-object collectives (gloo); halo
entries and dot products travel through peer memory inside libbtfem's kernels.  Rank 0 prints one JSON line:
DOF-steps/s of the partitioned solve (device time, max over ranks) and, with --check, the relative difference
of the signal to a single-GPU solve of the whole mesh on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nbox", type=int, default=78)
    ap.add_argument("--ecs", type=int, default=0, help="use the ECS slab with this many cells per edge instead")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--bvalue", type=float, default=1000.0)
    ap.add_argument("--trace", type=int, default=0, help="log the first N collective-closing kernels per rank")
    ap.add_argument("--comm", default="torch", choices=["torch", "socket"],
                    help="host plumbing of the set-up: torch.distributed (gloo) or the torch-free partition.SocketComm")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    entry.load_package()
    from dmri_fem_cloud_b200 import btfem, meshes, partition
    dist = None
    if args.comm == "socket":
        comm = partition.SocketComm.from_env()
    else:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        comm = partition.TorchComm(dist)

    Fb = pdir = None
    if args.ecs:
        # configs[3]: extracellular space of 226 cylinders, weak pseudo-periodic BC in x and y, g = (1,1,0)/sqrt2
        xyz, tets, phase = meshes.ecs_slab(args.ecs, args.ecs, 2)
        xyz, tets = meshes.coordinate_order(xyz, tets, axes=(1, 2, 0))       # vertex blocks = slabs normal to y
        mp, ts, f, fp = bench.sequence(delta=10000.0, Delta=13000.0, k=200.0, b=args.bvalue)
        tps = np.concatenate([[0.0], ts[:-1]])
        _, Fb = mp.profiles_on_grid(tps)
        g = np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
        pdir = [1, 1, 0]
        D, kappa, name = 2e-3, 1e-5, "configs[3] ECS slab %dx%dx2, 226 cylinders, weak periodic x,y" % (args.ecs, args.ecs)
    else:
        xyz, tets, phase = bench.workload(args.nbox)
        mp, ts, f, fp = bench.sequence(k=200.0, b=args.bvalue)
        g = np.array([0.0, 1.0, 0.0])
        D, kappa, name = 3e-3, 1e-5, "configs[1] cell-in-box n_box=%d" % args.nbox
    k, q = 200.0, mp.qvalue
    kw = dict(rtol=1e-9, atol=1e-10, maxit=100000)
    if pdir is not None:
        kw.update(q=q, Fb=Fb)
        lo, hi = xyz.min(axis=0), xyz.max(axis=0)

    t0 = time.perf_counter()
    d = partition.DistBTFem(xyz, tets, comm, device=local_rank, phase=phase)
    d.set_diffusion(D)
    d.set_relaxation(1e-16)
    d.set_permeability(kappa)
    if pdir is not None:
        hmin, _ = d.mesh_stats()
        d.set_periodic(pdir, 3e-3 / hmin, 1e-2 * hmin, lo, hi)        # kappa_e, tol as MyDomain sets them
    if args.trace:
        d.fem.dist_trace(args.trace)
    d.assemble()
    setup_s = time.perf_counter() - t0
    res = d.solve(k, 0.5, q * f, q * fp, g, **kw)            # warm-up
    if args.trace:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "trace_w%d_r%d.npy" % (world, rank)), d.fem.dist_get_trace())
    comm.barrier()
    loop_ms, wall = [], []
    for _ in range(args.reps):
        comm.barrier()
        t0 = time.perf_counter()
        res = d.solve(k, 0.5, q * f, q * fp, g, **kw)
        wall.append(time.perf_counter() - t0)
        loop_ms.append(res["loop_ms"] + res["setup_ms"])
    stats = comm.allgather(dict(rank=rank, loop_ms=min(loop_ms), wall_s=min(wall), n_own=d.n_own, n_int=d.n_int,
                                n_send=d.n_send, ndof_local=d.fem.ndof, nnz_local=d.fem.nnz, setup_s=setup_s))
    out = None
    if rank == 0:
        ndof_real = 2 * d.ndof_global
        tmax = max(s["loop_ms"] for s in stats) * 1e-3
        out = {"workload": name, "world": world, "ndof_real": ndof_real, "theta_steps": len(ts),
               "iters": int(res["total_iters"]), "loop_s_max_over_ranks": tmax,
               "wall_s_max_over_ranks": max(s["wall_s"] for s in stats),
               "dof_steps_per_s": ndof_real * len(ts) / tmax,
               "us_per_iteration": 1e6 * tmax / max(1, int(res["total_iters"])),
               "normalized_signal": res["signal"] / res["voi"], "kernels": int(res["n_kernels"]),
               "ranks": stats}
    d.close()
    if args.check and rank == 0:
        with btfem.BTFem(local_rank) as fem:
            fem.set_mesh(xyz, tets, phase)
            fem.set_diffusion(D)
            fem.set_relaxation(1e-16)
            fem.set_permeability(kappa)
            if pdir is not None:
                fem.set_periodic(pdir, 3e-3 / hmin, 1e-2 * hmin, lo, hi)
            fem.assemble()
            if pdir is not None:
                from dmri_fem_cloud_b200 import periodic
                dv, dc = fem.dofmap()
                fem.set_periodic_gather(*periodic.build_gather(xyz, tets, phase, pdir, lo, hi, dv, dc,
                                                               bfacets=fem.boundary_facets()))
            fem.solve(k, 0.5, q * f, q * fp, g, **kw)
            ref = fem.solve(k, 0.5, q * f, q * fp, g, **kw)
        out["single_gpu_loop_s"] = (ref["loop_ms"] + ref["setup_ms"]) * 1e-3
        out["single_gpu_iters"] = int(ref["total_iters"])
        out["single_gpu_signal"] = ref["signal"] / ref["voi"]
        out["rel_signal_err"] = abs(out["normalized_signal"] - out["single_gpu_signal"]) / abs(out["single_gpu_signal"])
        out["speedup_vs_single_gpu"] = out["single_gpu_loop_s"] / out["loop_s_max_over_ranks"]
    if rank == 0:
        print(json.dumps(out))
    comm.barrier()
    if dist is not None:
        dist.destroy_process_group()
    else:
        comm.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
