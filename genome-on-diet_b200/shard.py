"""Host-side sharding of a batch across the GPUs of one node: reads / pairs are independent
units, so rank i maps the i-th contiguous slice of the input with no data-path collective; what crosses ranks
is (a) the device-timed duration, reduced with MAX, and the unit counts, reduced with SUM (bench.py), and
(b) the per-shard results, gathered to rank 0 in rank order, which is input order (the reference's
kt_pipeline emits SAM records in input order, GDiet-ShortReads/kthread.c:101-115).

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """world+1 boundaries of contiguous, balanced shards of n units (the first n % world shards get one more)."""
    base, rem = divmod(int(n), int(world))
    sizes = np.full(world, base, np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def my_shard(n, rank, world):
    b = shard_bounds(n, world)
    return int(b[rank]), int(b[rank + 1])


def reduce_timing(ms, units, device="cpu", group=None):
    """(max over ranks of the per-rank times [list of floats], sum over ranks of the per-rank unit counts [list of ints])."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [float(x) for x in ms], [int(x) for x in units]
    t = torch.tensor([float(x) for x in ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    c = torch.tensor([int(x) for x in units], dtype=torch.int64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.tolist()], [int(x) for x in c.tolist()]


def gather_in_order(local, device="cpu", group=None):
    """Concatenate ragged 1-D per-rank arrays on rank 0 in rank (= input) order; other ranks get None.
    Returns (concatenated array, offsets[world+1]) on rank 0."""
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local, np.array([0, len(local)], np.int64)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    sizes[rank] = len(local)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
    sizes = sizes.cpu().numpy()
    cap = int(sizes.max()) if len(sizes) else 0
    buf = torch.zeros(cap, dtype=torch.from_numpy(local[:0]).dtype, device=device)
    buf[: len(local)] = torch.from_numpy(local).to(device)
    parts = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0, group=group)
    if rank != 0:
        return None, None
    out = np.concatenate([p.cpu().numpy()[: int(s)] for p, s in zip(parts, sizes)])
    return out, np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


class _DeviceBytes:
    """A raw device buffer exposed through __cuda_array_interface__ so that torch can wrap it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def broadcast_index(ctx, index, src=0, group=None, chunk_bytes=1 << 30):
    """The one collective of the mapper: rank `src` has built the index (gd_index_build); every other
    rank passes index=None and receives a replica in its own HBM.  The device buffers go over NCCL (NVLink /
    NVSwitch) straight from and into the library's allocations, in chunks of at most chunk_bytes."""
    import ctypes as C
    import gdiet_b200 as gd
    rank = dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    names = [f for f, _ in gd.gd_index_meta_t._fields_]
    meta_t = torch.zeros(len(names), dtype=torch.int64, device=dev)
    if rank == src:
        m = index.meta()
        meta_t = torch.tensor([int(getattr(m, f)) for f in names], dtype=torch.int64, device=dev)
    dist.broadcast(meta_t, src=src, group=group)
    if rank != src:
        m = gd.gd_index_meta_t(**{f: int(v) for f, v in zip(names, meta_t.tolist())})
        index = gd.Index.alloc(ctx, m)
    total = 0
    for ptr, nbytes in index.buffers():
        for o in range(0, nbytes, chunk_bytes):
            n = min(chunk_bytes, nbytes - o)
            t = torch.as_tensor(_DeviceBytes(ptr + o, n), device=dev)
            dist.broadcast(t, src=src, group=group)
            total += n
    torch.cuda.synchronize(dev)
    if rank != src:
        index.commit()
    return index, total
