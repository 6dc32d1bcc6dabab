"""N > 1 path on CPU: world_size-2 gloo processes each solve their contiguous shard of independent NMPC instances
(kernel bodies through the test-only host simulation: there is no GPU here) and rank 0 gathers u0 on the host —
exactly the structure bench.py uses under torchrun (no collective on the solve path)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.workloads import hostsim_model, make_rti_workload
from uclv_qs_pushing_matlab_b200 import sharding

B, N = 24, 10


def _solve(lo, hi):
    from tests.hostsim import hostsim as hs
    wl = make_rti_workload(None, batch=B, N=N, seed=2)
    mh = hostsim_model("santal")
    r = hs.solve([mh], N, 0.05, wl["x0"][lo:hi], wl["yref"][lo:hi], wl["yref_e"][lo:hi], np.zeros((hi - lo, N + 1, 4)),
                 wl["u_init"][lo:hi], mode="rti", prepare=True)
    return r["u"][:, 0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(B, world, rank)
    u0 = _solve(lo, hi)
    dist.barrier()
    full = sharding.gather_to_rank0(u0, B, world, rank)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _solve(0, B)
    assert full.shape == (B, 2) and np.array_equal(full, ref)
