"""CPU, world_size 2 over gloo: the N>1 plumbing of the sweep (sharding + signal gather)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402  (also runs in the spawned workers)

entry.load_package()
from dmri_fem_cloud_b200 import sweep  # noqa: E402


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10
    mine = sweep.shard_units(n, rank, world)
    sig = np.array([100.0 + u for u in mine])          # stands for this rank's normalized signals
    full = sweep.gather_signals(n, mine, sig, dist)
    dist.barrier()
    q.put((rank, full.tolist()))
    dist.destroy_process_group()


def test_two_rank_sweep_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [100.0 + u for u in range(10)]
    assert out[0] == want and out[1] == want
