"""CPU tests of the batched front end's host logic: contiguous pair sharding and the world-size-2
result gather (gloo), i.e. the N > 1 path of bench.py without a GPU."""
import os
import socket

import numpy as np
import pytest


def test_shard_range_partitions_contiguously():
    from opencv_contrib_b200.batch import shard_range
    for n, world in [(256, 8), (10, 3), (7, 8), (0, 4), (33, 2)]:
        seen = []
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            assert 0 <= lo <= hi <= n
            seen.extend(range(lo, hi))
        assert seen == list(range(n))          # every pair exactly once, in global order
        sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(256, 3, 8) == (96, 128)  # BASELINE configs[4]: 32 pairs per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opencv_contrib_b200.batch import gather_flows, shard_range
    n_pairs = 6
    lo, hi = shard_range(n_pairs, rank, world)
    # stand-in "flow fields": value encodes the global pair index
    local = torch.stack([torch.full((4, 5, 2), float(i)) for i in range(lo, hi)])
    out = gather_flows(local, dst=0)
    if rank == 0:
        allf = torch.cat(out, dim=0)
        q.put([float(allf[i, 0, 0, 0]) for i in range(n_pairs)])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gather_preserves_global_pair_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]
