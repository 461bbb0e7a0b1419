#!/usr/bin/env python
"""Multi-GPU short-read mapping, one process per GPU under torchrun:
rank 0 builds the device index once, its buffers are broadcast over NCCL (NVLink / NVSwitch) into every other GPU's
HBM (shard.broadcast_index -- the one collective of the path), then every rank maps its own contiguous shard of the
reads with no further communication.  Checks: every rank's replica answers lookups like rank 0's index; the
concatenation of the per-rank candidate lists in rank order equals what one GPU produces for the whole input.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_map.py"""
import json, os, sys, time, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import gdiet_b200 as gd
from gdiet_b200 import shard, synth


def main():
    ref_mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 50
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = gd.Context(local)
    genome = synth.random_genome(int(ref_mbp * 1e6), seed=1)
    reads = synth.sample_reads(genome, n_reads, 150, seed=2)       # every rank generates the same input, uses its shard
    opt = gd.sr_options()
    idx, t_build = None, 0.0
    if rank == 0:
        t0 = time.perf_counter()
        idx = ctx.index_build([genome], 11, 21, "10")
        t_build = time.perf_counter() - t0
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx, nbytes = shard.broadcast_index(ctx, idx, src=0)
    dist.barrier()
    t_bcast = time.perf_counter() - t0
    # replica check: a sample of lookups must agree with rank 0's answers
    probe = np.random.default_rng(5).integers(0, 1 << 42, 4096, dtype=np.uint64)
    keys = idx.export()[0]
    probe[: min(2048, len(keys))] = keys[:: max(1, len(keys) // 2048)][: min(2048, len(keys))]
    cnt, first = idx.get(probe)
    sig = torch.tensor([zlib.crc32(cnt.tobytes()), zlib.crc32(first.tobytes())], dtype=torch.int64, device=dev)
    sig0 = sig.clone()
    dist.broadcast(sig0, src=0)
    assert torch.equal(sig, sig0), "rank %d: replica lookups differ from rank 0" % rank
    lo, hi = shard.my_shard(n_reads, rank, world)
    off = np.arange(hi - lo, dtype=np.int64) * 150
    lens = np.full(hi - lo, 150, np.int32)
    buf = np.ascontiguousarray(reads[lo:hi].reshape(-1))
    ms = []
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, opt, cand_cap=(hi - lo) + 1024, cigar_cap=8 * (hi - lo) + 1024)
        ms.append((time.perf_counter() - t0) * 1e3)
    (t_map,), (n_cand, n_mapped) = shard.reduce_timing([ms[-1]], [len(cand), int((np.diff(coff) > 0).sum())], device=dev)
    # input-order gather of (rid, rs, score) per candidate and comparison with a single-GPU run on rank 0
    flat = np.stack([cand["rid"], cand["rs"], cand["score"], cand["n_cigar"]], 1).reshape(-1).astype(np.int32)
    allc, _ = shard.gather_in_order(flat, device=dev)
    out = None
    if rank == 0:
        off_all = np.arange(n_reads, dtype=np.int64) * 150
        lens_all = np.full(n_reads, 150, np.int32)
        coff1, cand1, _ = ctx.sr_map_batch(idx, off_all, lens_all, np.ascontiguousarray(reads.reshape(-1)), opt, cand_cap=n_reads + 1024,
                                           cigar_cap=8 * n_reads + 1024)
        flat1 = np.stack([cand1["rid"], cand1["rs"], cand1["score"], cand1["n_cigar"]], 1).reshape(-1).astype(np.int32)
        out = {"what": "multi-GPU sr mapping, index broadcast over NCCL", "n_gpus": world, "ref_bp": len(genome), "reads": n_reads,
               "index_build_s": round(t_build, 4), "index_bytes": nbytes, "broadcast_s": round(t_bcast, 4),
               "broadcast_gbs": nbytes / t_bcast / 1e9, "map_ms_max_over_ranks": round(t_map, 2), "reads_per_s_device_stage": n_reads / (t_map * 1e-3),
               "candidates": n_cand, "mapped_reads": n_mapped, "sharded_equals_single_gpu": bool(np.array_equal(allc, flat1))}
        print(json.dumps(out), flush=True)
    idx.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
