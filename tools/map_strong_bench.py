#!/usr/bin/env python
"""Strong-scaled short-read mapping on 1/2/4/8 GPUs (BASELINE.json north_star / config 5 shape; called by bench.py for
`extra.map_strong`, or stand-alone under torchrun):

  * a FIXED job -- 10 M synthetic 150 bp reads against a 3.1 Gbp (GRCh38-sized, 24 contigs) synthetic genome,
    `-ax sr -Z 10 -W 2 -k 21 -w 11 -r 0.05,150,200` -- whatever the number of GPUs;
  * rank 0 builds the sparsified index once on its GPU, the device buffers are broadcast over NCCL (NVLink / NVSwitch)
    into the other ranks' HBM (shard.broadcast_index: the one collective of the path);
  * every rank maps its own contiguous shard of the reads (pinned host ASCII in -> candidates + CIGARs out through
    gd_sr_map_batch, then the library's host stage -> SAM records), no communication;
  * the SAM text is gathered on rank 0 in rank order = input order and its SHA-256 is compared with the text rank 0
    produces alone for the whole input.

Times are max over ranks; reads/s = all reads / that time."""
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

READ_LEN = 150


def make_genome(dev, total_bp, ncontig=24, seed=6):
    import torch
    lens = np.full(ncontig, total_bp // ncontig, np.int32)
    off = np.zeros(ncontig, np.int64)
    off[1:] = np.cumsum(lens[:-1].astype(np.int64))
    total = int(lens.astype(np.int64).sum())
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    seq = torch.empty(total, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    step = 1 << 28
    for o in range(0, total, step):
        n = min(step, total - o)
        seq[o:o + n] = lut[torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev, generator=g).long()]
    return seq, off, lens


def make_reads(dev, genome, goff, glens, n_reads, seed=8, sub=0.01, indel_reads=0.25):
    """Illumina-like reads on the device (identical on every rank: same seed, same generator): uniform positions inside
    contigs, half reverse-complemented, 1 % substitutions, one 1-base insertion or deletion in a quarter of the reads."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    out = torch.empty((n_reads, READ_LEN), dtype=torch.uint8, device=dev)
    d_goff = torch.from_numpy(goff).to(dev)
    ar = torch.arange(READ_LEN, device=dev)
    clen = int(glens[0])
    step = 1 << 20
    for lo in range(0, n_reads, step):
        m = min(step, n_reads - lo)
        c = torch.randint(0, len(glens), (m,), device=dev, generator=g)
        p = torch.randint(0, clen - READ_LEN - 2, (m,), device=dev, generator=g) + d_goff[c]
        kind = torch.rand(m, device=dev, generator=g)            # < indel/2: deletion, < indel: insertion
        ip = torch.randint(5, READ_LEN - 5, (m,), device=dev, generator=g)
        is_del = (kind < indel_reads / 2)[:, None]
        is_ins = ((kind >= indel_reads / 2) & (kind < indel_reads))[:, None]
        src = ar[None, :] + (is_del & (ar[None, :] >= ip[:, None])).long() - (is_ins & (ar[None, :] > ip[:, None])).long()
        r = genome[(p[:, None] + src)]
        ins_base = lut[torch.randint(0, 4, (m,), device=dev, generator=g)]
        r = torch.where(is_ins & (ar[None, :] == ip[:, None]), ins_base[:, None], r)
        smask = torch.rand((m, READ_LEN), device=dev, generator=g) < sub
        sbase = lut[torch.randint(0, 4, (m, READ_LEN), device=dev, generator=g)]
        r = torch.where(smask, sbase, r)
        rev = torch.rand(m, device=dev, generator=g) < 0.5
        rc = comp[r.flip(1).long()]
        out[lo:lo + m] = torch.where(rev[:, None], rc, r)
    return out


def fixed_names(n):
    """'r' + 9 digits + NUL per read, and the pointer array gd_sr_sam_batch takes (built with numpy: no Python loop)."""
    idx = np.arange(n, dtype=np.int64)
    buf = np.zeros((n, 11), np.uint8)
    buf[:, 0] = ord("r")
    for d in range(9):
        buf[:, 9 - d] = ord("0") + (idx // 10 ** d) % 10
    ptr = (buf.ctypes.data + 11 * idx).astype(np.uint64)
    return buf, ptr


class StageFailed(RuntimeError):
    pass


def agree(dist, dev, err):
    """Every rank learns whether ANY rank failed the stage just finished, so that nobody waits in a collective for a rank
    that has already given up (err: None or a message)."""
    import torch
    bad = 1 if err else 0
    if dist:
        t = torch.tensor([bad], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        bad = int(t.item())
    if bad:
        raise StageFailed(err or "another rank failed")


def run(ctx, rank, world, dist, dev, n_reads=10_000_000, genome_bp=3_100_000_000, cores=None, log=None):
    import torch
    import gdiet_b200 as gd
    from gdiet_b200 import shard
    cores = cores or len(os.sched_getaffinity(0))
    L = ctx.lib
    t_all = time.perf_counter()
    err = None
    try:
        genome, goff, glens = make_genome(dev, genome_bp)
        total_bp = int(glens.astype(np.int64).sum())
        torch.cuda.synchronize(dev)
    except Exception as e:
        err = "genome: %s" % e
    agree(dist, dev, err)
    # ---- index: built once on rank 0, broadcast to the others --------------------------------------------------------
    idx, t_build = None, 0.0
    if rank == 0:
        try:
            t0 = time.perf_counter()
            idx = ctx.index_build_device(goff, glens, genome, 11, 21, "10")
            t_build = time.perf_counter() - t0
        except Exception as e:
            err = "index build: %s" % e
    agree(dist, dev, err)
    nbytes, t_bcast = 0, 0.0
    if world > 1:
        warm = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
        dist.broadcast(warm, src=0)  # communicator set-up is not the broadcast
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx, nbytes = shard.broadcast_index(ctx, idx, src=0)
        e1.record()
        e1.synchronize()
        (t_bcast,), _ = shard.reduce_timing([e0.elapsed_time(e1) * 1e-3], [0], device=dev)
    else:
        nbytes = sum(b for _, b in idx.buffers())
    # ---- reads (identical on every rank), this rank's shard in pinned host memory -------------------------------------
    try:
        return _map_stage(ctx, rank, world, dist, dev, n_reads, cores, L, idx, genome, goff, glens, total_bp, t_build, nbytes, t_bcast, t_all)
    finally:
        idx.close()


def _map_stage(ctx, rank, world, dist, dev, n_reads, cores, L, idx, genome, goff, glens, total_bp, t_build, nbytes, t_bcast, t_all):
    import torch
    import gdiet_b200 as gd
    from gdiet_b200 import shard
    err = None
    try:
        reads = make_reads(dev, genome, goff, glens, n_reads)
    except Exception as e:
        err = "reads: %s" % e
    agree(dist, dev, err)
    lo, hi = shard.my_shard(n_reads, rank, world)
    m = hi - lo
    h_reads = torch.empty((m, READ_LEN), dtype=torch.uint8).pin_memory()
    h_reads.copy_(reads[lo:hi])
    h_ref = genome.cpu().numpy()           # contigs as ASCII: the host stage reads them
    all_reads = reads.cpu().numpy() if rank == 0 and world > 1 else None
    del reads, genome
    torch.cuda.empty_cache()
    buf = h_reads.numpy().reshape(-1)
    off = np.arange(m, dtype=np.int64) * READ_LEN
    lens = np.full(m, READ_LEN, np.int32)
    qual = np.full(m * READ_LEN, ord("I"), np.uint8)
    name_buf, name_ptr = fixed_names(n_reads)
    seq_names = ["chr%d" % (i + 1) for i in range(len(glens))]
    opt = gd.sr_options()
    post = gd.sr_post_options(n_threads=max(1, cores // world))
    cand_cap, cig_cap = 2 * m + 1024, 12 * m + 1024
    h_cand = torch.zeros(cand_cap * 64, dtype=torch.uint8).pin_memory().numpy().view(gd.SR_CAND_DTYPE)
    h_cig = torch.zeros(cig_cap, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    h_coff = torch.zeros(m + 1, dtype=torch.int64).pin_memory().numpy()

    def map_call(ix, n, o, ln, b, coff, cand, cig):
        ncig = C.c_int64(0)
        rc = L.gd_sr_map_batch(ctx.h, ix.h, n, gd._ptr(o), gd._ptr(ln), gd._ptr(b), C.byref(opt), gd._ptr(coff), gd._ptr(cand), len(cand),
                               gd._ptr(cig), len(cig), C.byref(ncig))
        ctx._check(rc, "gd_sr_map_batch")
        return int(coff[n]), int(ncig.value)

    def sam_call(n, names, o, ln, b, q, coff, cand, cig, threads):
        post.n_threads = threads
        return gd.sr_sam_batch(names, o, ln, b, q, coff, cand, cig, seq_names, None, post, parts=True, ref=(goff, glens, h_ref))

    my_names = (C.c_char_p * m).from_buffer(name_ptr, lo * 8)
    # ---- e2e: reads in, SAM text out, the post-DP stage on the device (gd_sr_map_sam_batch) ---------------------------
    h_qual = torch.from_numpy(qual).pin_memory().numpy()
    t_dev, dev_txt = [], b""
    for it in range(4):  # two warm-up calls (the text lands in two alternating sets of pinned buffers, page-locked on first use), two timed
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)
        try:
            t0 = time.perf_counter()
            pieces = ctx.sr_map_sam_batch(idx, my_names, off, lens, buf, h_qual, opt, post, seq_names, join=False)
            if it >= 2:
                t_dev.append(time.perf_counter() - t0)
            if it == 3:
                dev_txt = b"".join(C.string_at(a, l) for a, l in pieces)
        except Exception as e:
            err = "map_sam: %s" % e
        agree(dist, dev, err)
    t_map, t_e2e, sam_txt, nc, ncg = [], [], b"", 0, 0
    for it in range(3):  # one warm-up, two timed
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)
        try:
            t0 = time.perf_counter()
            nc, ncg = map_call(idx, m, off, lens, buf, h_coff, h_cand, h_cig)
            t1 = time.perf_counter()
            parts = sam_call(m, my_names, off, lens, buf, qual, h_coff, h_cand[:max(nc, 1)], h_cig[:max(ncg, 1)], max(1, cores // world))
            t2 = time.perf_counter()
            if it:
                t_map.append(t1 - t0), t_e2e.append(t2 - t0)
            if it == 2:
                sam_txt = parts.bytes()
            parts.free()
        except Exception as e:
            err = "map: %s" % e
        agree(dist, dev, err)
    h2d = int(buf.nbytes + off.nbytes + lens.nbytes)
    d2h = int(nc * 64 + ncg * 4 + (m + 1) * 8)
    dev_same = int(dev_txt == sam_txt)
    (tm, te, td), (tot_cand, tot_cig, h2d, d2h, sam_bytes, dev_same) = shard.reduce_timing(
        [min(t_map), min(t_e2e), min(t_dev)], [nc, ncg, h2d, d2h, len(sam_txt), dev_same], device=dev if dist else "cpu")
    # ---- SAM records to rank 0 in input order ---------------------------------------------------------------------------
    t0 = time.perf_counter()
    gathered, _ = shard.gather_in_order(np.frombuffer(sam_txt, np.uint8), device=dev) if world > 1 else (np.frombuffer(sam_txt, np.uint8), None)
    t_gather = time.perf_counter() - t0
    out = None
    if rank == 0:
        sha = hashlib.sha256(gathered.tobytes()).hexdigest()
        same = None
        if world > 1:  # the same job on this GPU alone
            a_off = np.arange(n_reads, dtype=np.int64) * READ_LEN
            a_len = np.full(n_reads, READ_LEN, np.int32)
            a_buf = all_reads.reshape(-1)
            a_qual = np.full(n_reads * READ_LEN, ord("I"), np.uint8)
            a_coff = np.zeros(n_reads + 1, np.int64)
            a_cand = np.zeros(2 * n_reads + 1024, gd.SR_CAND_DTYPE)
            a_cig = np.zeros(12 * n_reads + 1024, np.uint32)
            nc1, ncg1 = map_call(idx, n_reads, a_off, a_len, a_buf, a_coff, a_cand, a_cig)
            p1 = sam_call(n_reads, (C.c_char_p * n_reads).from_buffer(name_ptr), a_off, a_len, a_buf, a_qual, a_coff, a_cand[:max(nc1, 1)],
                          a_cig[:max(ncg1, 1)], cores)
            same = hashlib.sha256(p1.bytes()).hexdigest() == sha
            p1.free()
        out = {"what": "strong scaling: %d x %d bp reads vs a %.2f Gbp synthetic genome (24 contigs), -ax sr -Z 10 -W 2 -k 21 -w 11 "
                       "-r 0.05,150,200; index built on rank 0 and broadcast over NCCL; contiguous read shards; SAM gathered in input order" % (
                           n_reads, READ_LEN, total_bp / 1e9),
               "n_gpus": world, "reads": n_reads, "genome_bp": total_bp, "host_cores": cores,
               "index_build_s": round(t_build, 4), "index_bytes": int(nbytes), "broadcast_s": round(t_bcast, 4),
               "broadcast_gbs": (nbytes / t_bcast / 1e9) if t_bcast > 0 else None,
               "map_call": {"s_max_over_ranks": round(tm, 4), "reads_per_s": n_reads / tm,
                            "note": "gd_sr_map_batch with pinned host buffers: H2D of the reads, all kernels, D2H of candidates + CIGARs"},
               "e2e": {"s_max_over_ranks": round(td, 4), "reads_per_s": n_reads / td, "h2d_bytes": int(h2d + n_reads * READ_LEN),
                       "d2h_bytes": sam_bytes, "identical_to_host_stage_on_ranks": dev_same,
                       "note": "gd_sr_map_sam_batch: pinned host reads + qualities in, SAM text out; the post-DP stage (mm_update_extra ... "
                               "mm_write_sam3) runs on the device, only text crosses PCIe"},
               "e2e_host_stage": {"s_max_over_ranks": round(te, 4), "reads_per_s": n_reads / te, "h2d_bytes": h2d, "d2h_bytes": d2h,
                                  "note": "map_call + the threaded HOST stage (gd_sr_sam_batch) on host_cores / n_gpus threads per rank"},
               "candidates": tot_cand, "cigar_entries": tot_cig, "sam_bytes": sam_bytes, "gather_s": round(t_gather, 4),
               "sam_sha256": sha, "sam_equals_single_gpu": same, "total_s": round(time.perf_counter() - t_all, 2)}
    return out


def main():
    import torch
    import gdiet_b200 as gd
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    gbp = float(sys.argv[2]) if len(sys.argv) > 2 else 3.1
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ctx = gd.Context(local)
    out = run(ctx, rank, world, dist, dev, n_reads, int(gbp * 1e9))
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
