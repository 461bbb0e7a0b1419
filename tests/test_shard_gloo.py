"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (contiguous input-order shards, MAX/SUM reduction of
the per-rank timings, input-order gather of ragged per-shard results).  No device code is involved."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import gdiet_b200  # noqa: F401
from gdiet_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.my_shard(n, rank, world)
        # every unit produces a ragged "CIGAR" of (i % 5) + 1 words holding its global index
        local = np.concatenate([np.full((i % 5) + 1, i, np.int32) for i in range(lo, hi)]) if hi > lo else np.zeros(0, np.int32)
        ms, units = shard.reduce_timing([10.0 + rank, 3.0 - rank], [hi - lo, len(local)])
        out, off = shard.gather_in_order(local)
        q.put((rank, lo, hi, ms, units, None if out is None else out.tolist(), None if off is None else off.tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            b = shard.shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0) and np.diff(b).max() - np.diff(b).min() <= 1


def test_two_rank_reduce_and_ordered_gather():
    world, n = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = np.concatenate([np.full((i % 5) + 1, i, np.int32) for i in range(n)])
    (r0, lo0, hi0, ms0, un0, out0, off0), (r1, lo1, hi1, ms1, un1, out1, off1) = res
    assert (lo0, hi1) == (0, n) and hi0 == lo1
    assert ms0 == ms1 == [11.0, 3.0]                       # MAX over ranks, per entry
    assert un0 == un1 == [n, len(exp)]                     # SUM over ranks
    assert out1 is None and np.array_equal(np.array(out0, np.int32), exp)  # input order on rank 0
    assert off0 == [0, int((np.arange(lo0, hi0) % 5 + 1).sum()), len(exp)]
