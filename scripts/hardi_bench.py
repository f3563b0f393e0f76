"""Config 5 (HARDI sweep on a neuron-like mesh, 46 k vertices ~ `25o_spindle17aFI`): 64 directions x 4 b-values,
sharded round-robin over ranks (torchrun), `batch` members per kernel launch.  Prints whole-job signals/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sympy as sp
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes, sweep

ndir = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

xyz, tets = meshes.neuron_like(h=0.7)
xyz, tets = meshes.rcm_order(*meshes.shuffle_vertices(xyz, tets, 0))
mp = dl.MRI_parameters()
mp.delta, mp.Delta = 10600.0, 43100.0
mp.T = mp.delta + mp.Delta
mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
mp.bvalue = 1000.0
mp.Apply()
sim = dl.MRI_simulation()
sim.k = 200.0
dirs = meshes.fibonacci_hemisphere(ndir)
bvals = [1000.0, 2000.0, 3000.0, 4000.0]
concurrent = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # > 0: that many handles, each on its share of the SMs
par = dict(rtol=1e-9, atol=1e-10, maxit=100000)


def make_fem(fem):
    fem.set_mesh(xyz, tets)
    fem.set_diffusion(3e-3)
    fem.set_relaxation(1e-16)


make_fem.device = local_rank
if concurrent > 0:
    fems = sweep.make_concurrent_handles(make_fem, concurrent)
    sweep.run_sweep_concurrent(fems, mp, sim, dirs[:concurrent // 2 + 1], bvals[:2], par)      # warm-up
else:
    fem = btfem.BTFem(local_rank)
    make_fem(fem)
    fem.assemble()
    sweep.run_sweep(fem, mp, sim, dirs[:2], bvals[:2], par, batch=min(batch, 4))          # warm-up
if dist is not None:
    dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
if concurrent > 0:
    mine, sig = sweep.run_sweep_concurrent(fems, mp, sim, dirs, bvals, par, rank=rank, world=world)
else:
    mine, sig = sweep.run_sweep(fem, mp, sim, dirs, bvals, par, rank=rank, world=world, batch=batch)
full = sweep.gather_signals(len(dirs) * len(bvals), mine, sig, dist)
if dist is not None:
    dist.barrier(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
if rank == 0:
    print("HARDI mesh %d verts %d tets | %d signals on %d GPU(s), batch %d: %.3f s -> %.2f signals/s | first %s | checksum %.12f"
          % (len(xyz), len(tets), len(full), world, batch if concurrent == 0 else -concurrent, dt, len(full) / dt,
             np.round(full[:4], 6), full.sum()))
if concurrent > 0 and rank == 0:
    st = sweep.run_sweep_concurrent.last_stats
    print("  per solve: loop %.1f ms, set-up %.1f ms, %.0f iterations -> %.1f us per iteration inside the kernel" % (
        st[:, 0].mean(), st[:, 1].mean(), st[:, 2].mean(), 1e3 * st[:, 0].sum() / st[:, 2].sum()))
if dist is not None:
    dist.destroy_process_group()
