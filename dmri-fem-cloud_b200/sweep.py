"""Direction x b-value sweeps (HARDI): the reference just loops serially over b-values / directions
(ExplicitImplementation.ipynb cell 10, Manifolds.ipynb cell 10).  Every (direction, b) solve is
independent, so the units shard over ranks (one process per GPU: directions round-robin, all b-values of a
direction on the same rank) with no data-path collective; the only exchange is the final gather of the signals."""
import numpy as np


def sweep_units(directions, bvalues):
    """All (direction index, b index) pairs, direction-major."""
    return [(i, j) for i in range(len(directions)) for j in range(len(bvalues))]


def shard_units(n_units, rank, world):
    """Round-robin ownership: unit u belongs to rank u % world."""
    return list(range(rank, n_units, world))


def shard_balanced(n_dir, n_b, rank, world):
    """This rank's units for a batched sweep: directions round-robin over ranks, every rank taking ALL b-values of
    its directions (the iteration count of a solve grows with b, so every rank gets the same mix), listed
    direction-major like sweep_units (unit id = i * n_b + j).  Measured on B200: batches that mix b-values are
    faster than b-homogeneous ones -- members that converge early drop out of the lock-step launches."""
    return [i * n_b + j for i in range(rank, n_dir, world) for j in range(n_b)]


def run_sweep(fem, mri_para, sim, directions, bvalues, linsolver_params, rank=0, world=1, batch=1):
    """Solve this rank's share.  Returns (unit ids, normalized signals).  `fem` is an assembled
    btfem.BTFem, `mri_para` a dmrifemlib.MRI_parameters with fs_sym/T set (Apply() is re-run per b).
    batch > 1: that many units advance in lock step per kernel launch (btfem_solve_batch)."""
    units = sweep_units(directions, bvalues)
    mine = shard_balanced(len(directions), len(bvalues), rank, world)
    ts = sim.time_grid(mri_para)
    tps = np.concatenate([[0.0], ts[:-1]])
    out = []
    cache = {}

    # The time profiles f, F and int F^2 do not depend on b: evaluate them (sympy) ONCE per sweep; only
    # q = sqrt(b / int F^2) changes (MRI_parameters.convert_b2q, DmriFemLib.py:847-849).  Re-running Apply() per
    # b-value costs ~0.15 s each on the host -- 40 % of an 8-GPU HARDI sweep that takes 1.3 s of GPU time per rank.
    mri_para.bvalue, mri_para.gvalue = bvalues[0], None
    mri_para.Apply()
    f, _ = mri_para.profiles_on_grid(ts)
    fp, Fp = mri_para.profiles_on_grid(tps)

    def scalars(j):
        if j not in cache:
            mri_para.bvalue, mri_para.gvalue = bvalues[j], None
            cache[j] = (mri_para.convert_b2q(), f, fp, Fp)
        return cache[j]

    def unit_dir(i):
        g = np.asarray(directions[i], dtype=float)
        return g / np.linalg.norm(g)

    if batch <= 1:
        for u in mine:
            i, j = units[u]
            q, f, fp, Fp = scalars(j)
            res = fem.solve(sim.k, sim.theta, q * f, q * fp, unit_dir(i), q=q, Fb=Fp, **linsolver_params)
            out.append(res["signal"] / res["voi"])
    else:
        for k0 in range(0, len(mine), batch):
            members = []
            for u in mine[k0:k0 + batch]:
                i, j = units[u]
                q, f, fp, _ = scalars(j)
                members.append((q * f, q * fp, unit_dir(i)))
            for res in fem.solve_batch(sim.k, sim.theta, members, **linsolver_params):
                out.append(res["signal"] / res["voi"])
    return mine, np.array(out)


def make_concurrent_handles(make_fem, n_handles, n_sms=148):
    """`n_handles` assembled btfem.BTFem objects on ONE GPU, each confined to its share of the SMs
    (btfem_set_sm_partition): their persistent time-loop kernels run side by side, so that a sweep over a mesh too
    small to fill the device keeps every SM and the whole memory system busy.  make_fem(fem) sets mesh and coefficients
    on a fresh handle (not assemble)."""
    from . import btfem as _bt
    fems = []
    share = max(1, n_sms // n_handles)
    try:
        for _ in range(n_handles):
            fem = _bt.BTFem(make_fem.device if hasattr(make_fem, "device") else 0)
            fems.append(fem)
            fem.set_sm_partition(share)
            make_fem(fem)
            fem.assemble()
    except Exception:
        for fem in fems:
            fem.close()
        raise
    return fems


def run_sweep_concurrent(fems, mri_para, sim, directions, bvalues, linsolver_params, rank=0, world=1):
    """Like run_sweep, with the rank's units solved CONCURRENTLY: one host thread per handle of `fems`
    (make_concurrent_handles) takes the next unit from a shared queue (largest b first: those solves take the most
    iterations) and runs it as one persistent-kernel launch on the handle's SM share.  Which handle solves a unit does
    not change its bits (all handles have the same launch shape)."""
    import queue
    import threading
    units = sweep_units(directions, bvalues)
    mine = shard_balanced(len(directions), len(bvalues), rank, world)
    ts = sim.time_grid(mri_para)
    tps = np.concatenate([[0.0], ts[:-1]])
    mri_para.bvalue, mri_para.gvalue = bvalues[0], None
    mri_para.Apply()
    f, _ = mri_para.profiles_on_grid(ts)
    fp, Fp = mri_para.profiles_on_grid(tps)
    qs = []
    for b in bvalues:
        mri_para.bvalue, mri_para.gvalue = b, None
        qs.append(mri_para.convert_b2q())
    order = sorted(range(len(mine)), key=lambda k: (-bvalues[units[mine[k]][1]], k))
    todo = queue.Queue()
    for k in order:
        todo.put(k)
    out = np.zeros(len(mine))
    errors = []
    stats = []            # (loop ms, set-up ms, iterations) per solve, for diagnostics: run_sweep_concurrent.last_stats

    def worker(fem):
        while True:
            try:
                k = todo.get_nowait()
            except queue.Empty:
                return
            i, j = units[mine[k]]
            g = np.asarray(directions[i], dtype=float)
            try:
                res = fem.solve(sim.k, sim.theta, qs[j] * f, qs[j] * fp, g / np.linalg.norm(g), q=qs[j], Fb=Fp,
                                **linsolver_params)
                out[k] = res["signal"] / res["voi"]
                stats.append((res["loop_ms"], res["setup_ms"], res["total_iters"]))
            except Exception as exc:          # keep the other workers going; report below
                errors.append(exc)
                return

    threads = [threading.Thread(target=worker, args=(fem,)) for fem in fems]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    run_sweep_concurrent.last_stats = np.array(stats, dtype=float)
    return mine, out


def gather_signals(n_units, mine, signals, dist=None):
    """Assemble the full signal table on every rank.  dist: an initialised torch.distributed module
    (gloo or nccl), a partition.SocketComm / TorchComm / LocalComm object (no PyTorch needed for the first), or None for
    a single process."""
    full = np.zeros(n_units)
    full[mine] = signals
    if dist is not None and hasattr(dist, "sum") and hasattr(dist, "world"):      # a partition.*Comm object
        return np.asarray(dist.sum(full)) if dist.world > 1 else full
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.from_numpy(full)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)           # disjoint supports: the sum is the gather
        full = t.cpu().numpy()
    return full
