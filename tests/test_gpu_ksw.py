"""GPU: bit-exact parity of the CUDA DP (called through the C ABI) with the oracle, the committed golden
vectors of the reference, and size-independent properties at the BASELINE sizes."""
import os

import numpy as np
import pytest

import gdiet_b200 as gd
from gdiet_b200 import synth
from oraclelib import EXTZ_FIELDS
from helpers import pair, oracle_batch, assert_batch_equal, params, cigar_spans

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCORINGS = ["sr", "map-hifi", "map-ont"]


def test_golden_vectors_through_batch_abi(ctx):
    """every golden case (flags x scorings x bands, N, code 7, degenerate lengths), grouped by (flag, scoring)"""
    g = np.load(os.path.join(GOLD, "ksw_golden.npz"))
    meta = g["meta"]
    keys = sorted(set((int(m[0]), int(m[1])) for m in meta))
    for flag, sci in keys:
        idx = np.nonzero((meta[:, 0] == flag) & (meta[:, 1] == sci))[0]
        qs = [g["qbuf"][g["qoff"][i]:g["qoff"][i] + g["qlen"][i]] for i in idx]
        ts = [g["tbuf"][g["toff"][i]:g["toff"][i] + g["tlen"][i]] for i in idx]
        P = synth.pack_pairs(qs, ts)
        w = np.ascontiguousarray(meta[idx, 2], np.int32)
        for G in (0, 4, 32):
            ctx.set_option("ksw_group", G)
            ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                                params(synth.SCORING[SCORINGS[sci]], flag), w=w)
            for k, i in enumerate(idx):
                mine = [int(ez[k][f]) for f in EXTZ_FIELDS]
                assert mine == [int(x) for x in g["ez"][i]], "golden case %d flag %#x G %d" % (i, flag, G)
                exp = g["cigar"][g["cigar_off"][i]:g["cigar_off"][i + 1]]
                assert np.array_equal(cig[coff[k]:coff[k + 1]], exp), "golden case %d flag %#x G %d" % (i, flag, G)
    ctx.set_option("ksw_group", 0)


def test_lead64_golden_vectors(ctx):
    """reference vectors (ksw_extd2_avx512) whose walk reads the AVX-512 build's lead-in cells (off[r] rounded down to 64):
    through the batch ABI and through the drop-in symbol"""
    from helpers import lead64_golden_cases
    ni = EXTZ_FIELDS.index("n_cigar")
    redone = 0
    for i, (q, t, sc, flag, w, differ, ez_exp, cig_exp) in enumerate(lead64_golden_cases()):
        P = synth.pack_pairs([q, q], [t, t])
        ctx.set_option("ksw_group", (0, 4, 32)[i % 3])
        ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, flag), w_all=w)
        for k in range(2):
            assert [int(ez[k][f]) for f in EXTZ_FIELDS] == ez_exp, "case %d" % i
            assert np.array_equal(cig[coff[k]:coff[k + 1]], cig_exp), "case %d" % i
        redone += int(ez[0]["lead64"])
        assert not differ or int(ez[0]["lead64"]) == 1
        if i % 4 == 0:
            e1, c1 = gd.ksw_extd2(q, t, synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"],
                                  sc["end_bonus"], flag, entry="ksw_extd2_avx512")
            assert [e1[f] for f in EXTZ_FIELDS] == ez_exp and np.array_equal(c1, cig_exp), "drop-in case %d" % i
    ctx.set_option("ksw_group", 0)
    assert redone >= 40


def test_narrow_band_random_scoring_sweep(ctx, oracle):
    """bounded slice of tools/ksw_fuzz.py --narrow: bands 0..5, scorings with b > q + 2e, every gang size"""
    rng = np.random.default_rng(640)
    redone = pairs = 0
    for it in range(24):
        e, q = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        sc = dict(a=int(rng.integers(1, 5)), b=int(rng.integers(q + 2 * e + 1, 2 * (q + e) + 1)), q=q, e=e,
                  q2=int(rng.integers(6, 30)), e2=1, zdrop=int(rng.choice([400, 40])), end_bonus=int(rng.choice([0, 5])))
        P = synth.ragged_pairs(150, seed=int(rng.integers(1 << 30)), max_len=int(rng.choice([150, 300, 700])))
        w = rng.choice([0, 1, 2, 3, 5], P["n"]).astype(np.int32)
        flag = int(rng.choice([0x08, 0x00, 0x0a, 0x88, 0x40, 0x18, 0xc2]))
        ctx.set_option("ksw_group", int(rng.choice([0, 4, 8, 16, 32])))
        ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, flag), w=w)
        assert_batch_equal(ez, coff, cig, oracle_batch(oracle, P, w, sc, flag), what="round %d flag %#x" % (it, flag))
        redone += int(ez["lead64"].sum())
        pairs += P["n"]
    ctx.set_option("ksw_group", 0)
    assert pairs == 3600 and redone >= 1


def test_random_scoring_flags_bands_sweep(ctx, oracle):
    """bounded slice of tools/ksw_fuzz.py: scoring values, flags, bands, lengths and gang sizes drawn at random (the other
    parity tests fix them to the reference's presets); 40 rounds of 120 pairs"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    import ksw_fuzz as kf
    rng = np.random.default_rng(2026)
    for it in range(40):
        sc = kf.draw_scoring(rng)
        flag = int(rng.choice(kf.FLAGS))
        P = synth.ragged_pairs(120, seed=int(rng.integers(1 << 30)), max_len=int(rng.choice([8, 40, 150, 300, 700])))
        w = rng.choice([-1, 0, 1, 3, 5, 10, 20, 33, 37, 64, 100, 150, 400, 1000], P["n"]).astype(np.int32)
        ctx.set_option("ksw_group", int(rng.choice([0, 4, 8, 16, 32])))
        ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, flag), w=w)
        assert_batch_equal(ez, coff, cig if not (flag & 1) else None, oracle_batch(oracle, P, w, sc, flag), what="round %d flag %#x %s" % (it, flag, sc))
    ctx.set_option("ksw_group", 0)


@pytest.mark.parametrize("flag", [0x08, 0x00, 0x18, 0x40, 0x48, 0xc2, 0x01, 0x0a, 0x42, 0x80])
def test_ragged_pairs_vs_oracle(ctx, oracle, flag):
    P = synth.ragged_pairs(300, seed=100 + flag, max_len=260)
    rng = np.random.default_rng(flag)
    w = rng.choice([-1, 5, 10, 20, 33, 37, 100, 150, 400], P["n"]).astype(np.int32)
    for scn in ("sr", "map-ont"):
        sc = synth.SCORING[scn]
        exp = oracle_batch(oracle, P, w, sc, flag)
        for G in (4, 8, 16, 32):
            ctx.set_option("ksw_group", G)
            ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                                params(sc, flag), w=w)
            assert_batch_equal(ez, coff, cig if not (flag & 1) else None, exp, what="flag %#x G %d %s" % (flag, G, scn))
    ctx.set_option("ksw_group", 0)


def test_swapped_gap_pieces_and_qe_seed(ctx, oracle):
    """q2+e2 < q+e: the kernel must order the pieces but seed H with the caller's q+e (reference quirk)"""
    P = synth.ragged_pairs(120, seed=31, max_len=200)
    sc = dict(synth.SCORING["sr"])
    sc["q"], sc["e"], sc["q2"], sc["e2"] = sc["q2"], sc["e2"], sc["q"], sc["e"]
    w = np.full(P["n"], 100, np.int32)
    for flag in (0x08, 0x00):
        exp = oracle_batch(oracle, P, w, sc, flag)
        ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                            params(sc, flag), w=w)
        assert_batch_equal(ez, coff, cig, exp)


@pytest.mark.parametrize("flag", [0x18, 0x58])
def test_approx_drop_on_last_row(ctx, oracle, flag):
    """KSW_EZ_APPROX_DROP: a Z-drop on the last anti-diagonal leaves before ez->score is set (ksw2_extd2_sse.c:380-382)"""
    for max_len, zdrop in ((8, 0), (8, 5), (40, 5), (150, 40)):
        P = synth.ragged_pairs(400, seed=max_len + zdrop, max_len=max_len)
        w = np.random.default_rng(zdrop).choice([-1, 1, 2, 5, 33, 100], P["n"]).astype(np.int32)
        sc = dict(synth.SCORING["sr"], zdrop=zdrop)
        exp = oracle_batch(oracle, P, w, sc, flag)
        assert any(e[0]["zdropped"] and e[0]["score"] < -(1 << 29) for e in exp)
        for G in (0, 4, 32):
            ctx.set_option("ksw_group", G)
            ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                                params(sc, flag), w=w)
            assert_batch_equal(ez, coff, cig, exp, what="flag %#x G %d zdrop %d" % (flag, G, zdrop))
    ctx.set_option("ksw_group", 0)


def test_long_banded_pairs_ring_wrap(ctx, oracle):
    """HiFi/ONT-like shapes at a size the oracle finishes in seconds; the column ring wraps hundreds of times"""
    for n, qlen, edit, wv, scn in ((6, 3000, 0.08, 200, "map-ont"), (4, 4000, 0.01, 500, "map-hifi")):
        P = synth.long_pairs(n, qlen, edit, seed=qlen, tlen_extra=0.02)
        w = np.full(P["n"], wv, np.int32)
        sc = synth.SCORING[scn]
        for flag in (0x08, 0x00):
            exp = oracle_batch(oracle, P, w, sc, flag)
            for G in (0, 16):
                ctx.set_option("ksw_group", G)
                ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                                    params(sc, flag), w=w)
                assert_batch_equal(ez, coff, cig, exp, what="long %d flag %#x" % (qlen, flag))
    ctx.set_option("ksw_group", 0)


def test_baseline_long_read_shapes(ctx, oracle):
    """BASELINE configs 3 and 4 at full size: 15 kbp HiFi-like pairs with -r 1000 and 50 kbp ONT-like pairs with
    -r 1300 (live flag; HiFi also with the exact maximum), sequences read from global memory, ring of > 1000 columns"""
    for n, qlen, edit, wv, scn, flags in ((2, 15000, 0.01, 1000, "map-hifi", (0x08, 0x00)), (2, 50000, 0.08, 1300, "map-ont", (0x08,))):
        P = synth.long_pairs(n, qlen, edit, seed=qlen + 1, tlen_extra=0.01)
        w = np.full(P["n"], wv, np.int32)
        sc = synth.SCORING[scn]
        for flag in flags:
            exp = oracle_batch(oracle, P, w, sc, flag)
            ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                                params(sc, flag), w=w)
            assert_batch_equal(ez, coff, cig, exp, what="baseline long %d flag %#x" % (qlen, flag))


def test_chunked_backtrack_arena(ctx, oracle):
    """a tiny backtrack budget forces many chunks; results must not depend on chunking"""
    P = synth.ksw_pairs(500, 150, 200, 0.05, seed=3)
    w = np.full(P["n"], 150, np.int32)
    sc = synth.SCORING["sr"]
    ctx.set_option("p_budget_mb", 2)
    ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, 0x08), w=w)
    assert ctx.stat("ksw_chunks") > 5
    ctx.set_option("p_budget_mb", 0)
    ez2, coff2, cig2 = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, 0x08), w=w)
    assert np.array_equal(ez, ez2) and np.array_equal(coff, coff2) and np.array_equal(cig, cig2)
    idx = list(range(0, 500, 25))
    assert_batch_equal(ez, coff, cig, oracle_batch(oracle, P, w, sc, 0x08, idx), idx)


def test_dropin_symbols_match_oracle(oracle):
    """ksw_extd2_avx512 / ksw_extd2_sse with the reference's own signature and ksw_extz_t"""
    P = synth.ragged_pairs(40, seed=77, max_len=220)
    sc = synth.SCORING["sr"]
    mat = synth.score_matrix(sc["a"], sc["b"])
    for i in range(P["n"]):
        q, t = pair(P, i)
        flag = [0x08, 0x00, 0x40, 0x01][i % 4]
        w = [-1, 20, 150][i % 3]
        exp = oracle.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag)
        for entry in ("ksw_extd2_avx512", "ksw_extd2_sse"):
            ez, cig = gd.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag,
                                   entry=entry)
            assert ez == exp[0] and np.array_equal(cig, exp[1]), (i, entry)


def test_dropin_thread_contexts_are_pooled(oracle):
    """kt_for starts fresh threads for every mini-batch (kthread.c:54-69): the drop-in symbols must not leave one context
    (streams, events, device and pinned buffers) behind per thread ever started"""
    import threading
    L = gd.load()
    P = synth.ragged_pairs(8, seed=5, max_len=120)
    sc = synth.SCORING["sr"]
    mat = synth.score_matrix(sc["a"], sc["b"])
    bad = []

    def work(i):
        q, t = pair(P, i % P["n"])
        exp = oracle.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], 50, sc["zdrop"], sc["end_bonus"], 0x08)
        ez, cig = gd.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], 50, sc["zdrop"], sc["end_bonus"], 0x08)
        if ez != exp[0] or not np.array_equal(cig, exp[1]):
            bad.append(i)

    for batch in range(6):  # six "mini-batches" of three fresh worker threads each
        ths = [threading.Thread(target=work, args=(3 * batch + k,)) for k in range(3)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        assert 1 <= L.gd_thread_ctx_pool_size(0) <= 3, "18 threads so far must have shared at most 3 contexts"
    assert not bad


def test_empty_and_degenerate_inputs(ctx):
    sc = synth.SCORING["sr"]
    z = np.zeros(0, np.int32)
    ez, coff, cig = ctx.ksw_extd2_batch(z, np.zeros(0, np.int64), np.zeros(1, np.uint8), z, np.zeros(0, np.int64),
                                        np.zeros(1, np.uint8), params(sc, 0x08))
    assert len(ez) == 0 and coff[0] == 0
    # qlen == 0 / tlen == 0 pairs return the reset ksw_extz_t (ksw2_extd2_sse.c:75-76)
    P = synth.pack_pairs([np.zeros(0, np.uint8), np.array([1, 2], np.uint8)], [np.array([1], np.uint8), np.zeros(0, np.uint8)])
    ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], np.append(P["qbuf"], 0).astype(np.uint8), P["tlen"], P["toff"],
                                        np.append(P["tbuf"], 0).astype(np.uint8), params(sc, 0x08), w_all=10)
    for i in range(2):
        assert ez[i]["score"] == gd.KSW_NEG_INF and ez[i]["n_cigar"] == 0 and ez[i]["max"] == 0 and ez[i]["zdropped"] == 0
    # mismatch penalty beyond 2(q+e): the reference returns before the DP (ksw2_extd2_sse.c:100)
    P = synth.ksw_pairs(4, 50, 60, 0.05, seed=1)
    prm = gd.KswParams(synth.score_matrix(2, 60), 12, 2, 24, 1, 100, 10, 0x08)
    ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], prm, w_all=50)
    assert all(int(e["score"]) == gd.KSW_NEG_INF and int(e["n_cigar"]) == 0 for e in ez)
    with pytest.raises(gd.GdietError):  # unsupported flag is an error, not a silent different answer
        ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, 0x04), w_all=50)


def test_baseline_size_properties(ctx, oracle):
    """Config 2 shape at 200k pairs: (1) every CIGAR spans exactly the aligned query/target ranges
    (the assert of mm_update_extra, GDiet-ShortReads/align.c:314); (2) identical pairs -> score = a*len,
    one M op; (3) results independent of batch order (shuffle -> same per-pair records);
    (4) a strided sample equals the oracle."""
    n = 200_000
    P = synth.ksw_pairs_fast(n, 150, 200, 0.05, seed=3)
    sc = synth.SCORING["sr"]
    w = np.full(n, 150, np.int32)
    for flag in (0x08, 0x00):
        ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"], params(sc, flag), w=w)
        ok = ez["zdropped"] == 0
        assert ok.mean() > 0.5
        ops = cig & 0xf
        lens = (cig >> 4).astype(np.int64)
        pid = np.repeat(np.arange(n), np.diff(coff))
        qspan = np.bincount(pid, weights=np.where((ops == 0) | (ops == 1), lens, 0), minlength=n).astype(np.int64)
        tspan = np.bincount(pid, weights=np.where((ops == 0) | (ops == 2), lens, 0), minlength=n).astype(np.int64)
        assert np.all(ez["n_cigar"] == np.diff(coff))
        assert np.all(qspan[ok] == 150) and np.all(tspan[ok] == 200)
        idx = list(range(0, n, n // 64))
        assert_batch_equal(ez, coff, cig, oracle_batch(oracle, P, w, sc, flag, idx), idx, what="sample flag %#x" % flag)
        # order independence
        perm = np.random.default_rng(1).permutation(n)[:20000]
        ez2, coff2, cig2 = ctx.ksw_extd2_batch(P["qlen"][perm], P["qoff"][perm], P["qbuf"], P["tlen"][perm], P["toff"][perm],
                                               P["tbuf"], params(sc, flag), w=w[perm])
        assert np.array_equal(ez2, ez[perm])
        for k in range(0, 20000, 997):
            assert np.array_equal(cig2[coff2[k]:coff2[k + 1]], cig[coff[perm[k]]:coff[perm[k] + 1]])
    # identical sequences
    t = synth.random_codes(np.random.default_rng(5), 150 * 1000).reshape(1000, 150)
    Pi = synth.pack_pairs(list(t), list(t))
    ez, coff, cig = ctx.ksw_extd2_batch(Pi["qlen"], Pi["qoff"], Pi["qbuf"], Pi["tlen"], Pi["toff"], Pi["tbuf"], params(sc, 0x08), w_all=150)
    assert np.all(ez["score"] == 300) and np.all(ez["n_cigar"] == 1) and np.all(cig == (150 << 4))
