"""GPU: bit-exact parity of the CUDA sketching with the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest

import gdiet_b200 as gd
from gdiet_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _xy(a):
    return np.stack([a["x"], a["y"]], 1) if len(a) else np.zeros((0, 2), np.uint64)


def test_golden_vectors_through_dropin_symbols():
    g = np.load(os.path.join(GOLD, "sketch_golden.npz"))
    pats = [str(p) for p in g["patterns"]]
    seq_all = g["seq"].tobytes()
    for i in range(len(g["meta"])):
        kind, k, w, zi, fn, a0, a1 = (int(x) for x in g["meta"][i])
        seq = seq_all[g["seq_off"][i]:g["seq_off"][i + 1]]
        exp = g["entries"][g["entries_off"][i]:g["entries_off"][i + 1]]
        extra = g["extra"][g["extra_off"][i]:g["extra_off"][i + 1]]
        Z = pats[zi]
        if fn == 0:
            got = gd.mm_sketch(seq, w, k, a0, Z)
        elif fn == 1:
            got, ret = gd.mm_sketch3(seq, w, k, 0, Z, a0, a1)
            assert ret == int(extra[0]), "case %d ret" % i
        else:
            got, counts = gd.mm_sketch2(seq, w, k, 0, Z, a0 / 1000.0)
            assert np.array_equal(counts, extra), "case %d counts" % i
        assert np.array_equal(got, exp), "case %d (fn %d kind %d k %d w %d Z %s)" % (i, fn, kind, k, w, Z)


def test_ref_batch_vs_oracle_multi_tile(ctx, oracle):
    """contigs of very different lengths (multi-tile path, tile seams), with N runs; rid stamped"""
    rng = np.random.default_rng(4)
    lens = [1, 37, 150, 2047, 2048, 2049, 5000, 40000, 250000]
    seqs = []
    for n in lens:
        c = rng.integers(0, 4, n)
        if n > 1000:
            for _ in range(4):
                a = int(rng.integers(0, n))
                c[a:a + int(rng.integers(1, 60))] = 4
        seqs.append(bytes(synth.ACGTN[c]))
    buf = np.frombuffer(b"".join(seqs), np.uint8)
    ln = np.array(lens, np.int32)
    off = np.zeros(len(lens), np.int64)
    off[1:] = np.cumsum(ln[:-1])
    rid = np.arange(len(lens), dtype=np.uint32) + 3
    for (k, w, Z) in ((21, 11, "10"), (19, 19, "10"), (15, 10, "110"), (28, 30, "101001"), (12, 200, "1")):
        oo, out = ctx.sketch_ref_batch(off, ln, buf, w, k, Z, rid=rid)
        for i, s in enumerate(seqs):
            exp = oracle.mm_sketch(s, w, k, int(rid[i]), Z)
            assert np.array_equal(_xy(out[oo[i]:oo[i + 1]]), exp), (k, w, Z, i)


def test_reads_batch_vs_oracle(ctx, oracle):
    """the per-read pair of calls of GDiet-ShortReads/map.c:74-99 for a whole batch"""
    g = synth.random_genome(200000, seed=9)
    for (k, w, Z, rl, ms, cap) in ((21, 11, "10", 150, 0.1, 800), (15, 10, "10", 3000, 0.1, 0xffffffff),
                                   (19, 19, "110", 1500, 2.0, 5), (21, 11, "100", 150, 0.5, 3)):
        reads = synth.sample_reads(g, 64, rl, seed=k)
        if k == 15:
            reads[3, 100:130] = ord("N")
        n = len(reads)
        buf = np.ascontiguousarray(reads.reshape(-1))
        ln = np.full(n, rl, np.int32)
        off = np.arange(n, dtype=np.int64) * rl
        R = ctx.sketch_reads_batch(off, ln, buf, w, k, Z, ms, cap)
        W = len(Z)
        for i in range(n):
            s = reads[i].tobytes()
            e2, c2 = oracle.mm_sketch2(s, w, k, 0, Z, ms)
            assert np.array_equal(R["s2_counts"][i], c2), (k, i)
            assert np.array_equal(_xy(R["s2"][R["s2_off"][i]:R["s2_off"][i + 1]]), e2), (k, i)
            for sh in range(W):
                e3, ret = oracle.mm_sketch3(s, w, k, 0, Z, sh, cap)
                q = i * W + sh
                assert np.array_equal(_xy(R["s3"][R["s3_off"][q]:R["s3_off"][q + 1]]), e3), (k, i, sh)
                assert int(R["s3_ret"][i, sh]) == ret, (k, i, sh)


def test_full_size_properties(ctx, oracle):
    """20 Mbp contig (config-5 shape, reduced): output sorted by position, unique, density ~ 2/(w+1) per
    sparsified base, chunk-independent (two halves with overlap reproduce the whole), and a window of the
    sequence re-sketched by the oracle agrees record for record."""
    n = 20_000_000
    gnm = synth.random_genome(n, seed=6)
    k, w, Z = 21, 11, "10"
    oo, out = ctx.sketch_ref_batch(np.zeros(1, np.int64), np.array([n], np.int32), gnm, w, k, Z, rid=np.array([5], np.uint32))
    pos = (out["y"] & 0xffffffff) >> 1
    assert np.all(np.diff(pos.astype(np.int64)) > 0)
    assert np.all((out["y"] >> 32) == 5) and np.all((out["x"] & 0xff) == k)
    dens = len(out) / (n / 2)
    assert abs(dens - 2.0 / (w + 1)) < 0.01
    # a 30 kbp window: records strictly inside (away from the window edges) must agree with the oracle
    a = 7_000_000
    sub = gnm[a:a + 30000].tobytes()
    exp = oracle.mm_sketch(sub, w, k, 5, Z)
    epos = ((exp[:, 1] & 0xffffffff) >> 1).astype(np.int64) + a
    inner = (epos > a + 200) & (epos < a + 30000 - 200)
    sel = (pos.astype(np.int64) > a + 200) & (pos.astype(np.int64) < a + 30000 - 200)
    assert np.array_equal(pos[sel].astype(np.int64), epos[inner])
    assert np.array_equal(out["x"][sel], exp[inner, 0])
