"""GPU, two devices: gd_multi_* (include/gdiet_cuda.h section 5) -- index built once and broadcast (NCCL, or peer copies),
contiguous read shards on every device, results in input order equal to a single device's."""
import numpy as np
import pytest

import gdiet_b200 as gd
import maplib

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def flat(reads):
    lens = np.array([len(r) for r in reads], np.int32)
    off = np.zeros(len(reads), np.int64)
    off[1:] = np.cumsum(lens[:-1].astype(np.int64))
    return off, lens, np.concatenate(reads)


@pytest.mark.parametrize("n_dev", [1, 2])
def test_multi_short_reads_equal_single_device(ctx, n_dev, monkeypatch):
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    contigs, reads = maplib.make_dataset(seed=71, n_reads=5000)
    rng = np.random.default_rng(2)
    reads = [r[:int(rng.integers(80, 151))].copy() for r in reads]  # ragged: shards are cut by bases, not by reads
    off, lens, buf = flat(reads)
    o = maplib.sr_opt(min_cnt=0.2, rec_frac=0.1)
    idx1 = ctx.index_build(contigs, 11, 21, "10")
    want = ctx.sr_map_batch(idx1, off, lens, buf, o)
    names = ["r%d" % i for i in range(len(reads))]
    qual = np.full(len(buf), 70, np.uint8)
    seq_names = ["chr%d" % (i + 1) for i in range(len(contigs))]
    post = gd.sr_post_options(n_threads=4)
    want_sam = gd.sr_sam_batch(names, off, lens, buf, qual, want[0], want[1], want[2], seq_names, contigs, post)
    for no_nccl in (False, True):
        if no_nccl:
            monkeypatch.setenv("GDIET_NO_NCCL", "1")
        M = gd.Multi(n_dev)
        st = M.index_bcast(M.ctx(0).index_build(contigs, 11, 21, "10"))
        if n_dev > 1:
            assert st["bcast_path"] == (2 if no_nccl else 1) and st["bcast_bytes"] > 0
            k1, c1, p1, s1 = idx1.export()
            k2, c2, p2, s2 = M.index(1).export()  # the replica answers like the index it was copied from
            assert np.array_equal(k1, k2) and np.array_equal(c1, c2) and np.array_equal(p1, p2) and np.array_equal(s1, s2)
        got = M.map_batch(off, lens, buf, o)
        assert np.array_equal(got[0], want[0])
        for f in maplib.CAND_FIELDS:
            assert np.array_equal(got[1][f], want[1][f]), f
        assert np.array_equal(got[2], want[2])
        sam = M.map_sam(names, off, lens, buf, qual, o, post, seq_names, contigs)
        assert sam == bytes(want_sam)
        M.close()
    idx1.close()


def test_multi_long_reads_equal_single_device(ctx):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    contigs, reads = maplib.make_long_dataset(seed=72, read_len=7000, n_reads=48)
    off, lens, buf = flat(reads)
    idx1 = ctx.index_build(contigs, 19, 19, "10")
    o = gd.lr_options("map-hifi", bw=800, mid_occ=50)
    want = ctx.lr_map_batch(idx1, off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=1 << 22)
    M = gd.Multi(2)
    M.index_bcast(M.ctx(0).index_build(contigs, 19, 19, "10"))
    got = M.map_batch(off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=1 << 22)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
    for f in maplib.CAND_FIELDS:
        assert np.array_equal(got[1][f], want[1][f]), f
    assert np.array_equal(got[1]["reserved"], want[1]["reserved"])
    M.close()
    idx1.close()


def test_multi_long_read_sam_in_slices_equals_one_batch(ctx, monkeypatch):
    """gd_multi_lr_map_sam cuts a long-read shard into device slices whose host SAM stage overlaps the next slice's mapping;
    the text of 6 slices equals the single-batch text (device stage + gd_lr_sam_batch on the whole batch)."""
    contigs, reads = maplib.make_long_dataset(seed=73, read_len=6000, n_reads=50, sv_frac=0.5)
    off, lens, buf = flat(reads)
    qual = np.full(len(buf), ord("I"), np.uint8)
    names = ["r%d" % i for i in range(len(reads))]
    seq_names = ["chr%d" % (i + 1) for i in range(len(contigs))]
    o = gd.lr_options("map-hifi", bw=600, mid_occ=50)
    post = gd.lr_post_options("map-hifi")
    idx1 = ctx.index_build(contigs, 19, 19, "10")
    coff, cand, cig = ctx.lr_map_batch(idx1, off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=1 << 22)
    want, _, _ = gd.lr_sam_batch(names, off, lens, buf, qual, coff, cand, cig, seq_names, contigs, post)
    assert len(want) > 50 * 6000
    M = gd.Multi(1)
    M.index_bcast(M.ctx(0).index_build(contigs, 19, 19, "10"))
    whole = M.map_sam(names, off, lens, buf, qual, o, post, seq_names, contigs)
    monkeypatch.setenv("GDIET_LR_SLICE_READS", "8")
    monkeypatch.setenv("GDIET_LR_SLICE_BASES", "0")
    sliced = M.map_sam(names, off, lens, buf, qual, o, post, seq_names, contigs)
    assert whole == bytes(want) and sliced == bytes(want)
    M.close()
    idx1.close()
