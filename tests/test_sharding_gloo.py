"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch by code block, 'decode' their shard (the oracle
stands in for the GPU here - test infrastructure), gather, and must reproduce the single-process result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_balance():
    sys.path.insert(0, ROOT)
    from srsran_4g_b200.sharding import shard_bounds
    for world in (1, 2, 4, 8):
        b = shard_bounds(np.full(16384, 6144), world)
        assert b[0][0] == 0 and b[-1][1] == 16384 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    rng = np.random.default_rng(0)
    K = rng.choice([40, 512, 1024, 6144], 1000)
    for world in (2, 4, 8):
        b = shard_bounds(K, world)
        sums = [K[lo:hi].sum() for lo, hi in b]
        assert b[-1][1] == 1000 and max(sums) - min(sums) <= 2 * 6144
    assert shard_bounds([], 4) == [(0, 0)] * 4
    assert shard_bounds([40], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as ol
    import vecgen
    from srsran_4g_b200.sharding import gather_results, shard_bounds
    K, n = 512, 21
    _, llr = vecgen.make_cb_batch(K, n, 1.5, 9)
    bounds = shard_bounds(np.full(n, K), world)
    lo, hi = bounds[rank]
    _, out, noi, ok = ol.oracle().tdec_batch(K, llr[lo:hi], 8, True)
    full = gather_results(out, noi, ok, bounds, dist)
    if rank == 0:
        _, out1, noi1, ok1 = ol.oracle().tdec_batch(K, llr, 8, True)
        q.put(bool((full[0] == out1).all() and (full[1] == noi1).all() and (full[2] == ok1).all()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
