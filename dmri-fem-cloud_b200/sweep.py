"""Direction x b-value sweeps (HARDI): the reference just loops serially over b-values / directions
(ExplicitImplementation.ipynb cell 10, Manifolds.ipynb cell 10).  Every (direction, b) solve is
independent, so the units shard round-robin over ranks (one process per GPU) with no data-path
collective; the only exchange is the final gather of the signals."""
import numpy as np


def sweep_units(directions, bvalues):
    """All (direction index, b index) pairs, direction-major."""
    return [(i, j) for i in range(len(directions)) for j in range(len(bvalues))]


def shard_units(n_units, rank, world):
    """Round-robin ownership: unit u belongs to rank u % world."""
    return list(range(rank, n_units, world))


def run_sweep(fem, mri_para, sim, directions, bvalues, linsolver_params, rank=0, world=1):
    """Solve this rank's share.  Returns (unit ids, normalized signals).  `fem` is an assembled
    btfem.BTFem, `mri_para` a dmrifemlib.MRI_parameters with fs_sym/T set (Apply() is re-run per b)."""
    units = sweep_units(directions, bvalues)
    mine = shard_units(len(units), rank, world)
    ts = sim.time_grid(mri_para)
    tps = np.concatenate([[0.0], ts[:-1]])
    out = []
    cache = {}
    for u in mine:
        i, j = units[u]
        if j not in cache:
            mri_para.bvalue, mri_para.gvalue = bvalues[j], None
            mri_para.Apply()
            f, _ = mri_para.profiles_on_grid(ts)
            fp, Fp = mri_para.profiles_on_grid(tps)
            cache[j] = (mri_para.qvalue, f, fp, Fp)
        q, f, fp, Fp = cache[j]
        g = np.asarray(directions[i], dtype=float)
        g = g / np.linalg.norm(g)
        res = fem.solve(sim.k, sim.theta, q * f, q * fp, g, q=q, Fb=Fp, **linsolver_params)
        out.append(res["signal"] / res["voi"])
    return mine, np.array(out)


def gather_signals(n_units, mine, signals, dist=None):
    """Assemble the full signal table on every rank.  dist: an initialised torch.distributed module
    (gloo or nccl) or None for a single process."""
    full = np.zeros(n_units)
    full[mine] = signals
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.from_numpy(full)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)           # disjoint supports: the sum is the gather
        full = t.cpu().numpy()
    return full
