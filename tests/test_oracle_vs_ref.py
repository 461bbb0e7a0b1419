"""CPU: pin the oracle against the UNMODIFIED reference compiled into oracle/_ref (skipped where the
reference objects are absent and the host cannot build them)."""
import numpy as np

import gdiet_b200  # noqa: F401
from gdiet_b200 import synth
from helpers import pair

FLAGS = [0x08, 0x00, 0x18, 0x40, 0x48, 0xc2, 0x01, 0x0a, 0x42, 0x80]


def _same(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1])


def test_ksw_oracle_vs_reference_sse_and_avx512(oracle, ref_avx):
    rng = np.random.default_rng(12)
    P = synth.ragged_pairs(600, seed=21, max_len=260)
    names = list(synth.SCORING)
    for i in range(P["n"]):
        q, t = pair(P, i)
        sc = dict(synth.SCORING[names[i % 3]])
        if i % 11 == 0:  # second piece cheaper than the first: exercises the swap and the qe seed quirk
            sc["q"], sc["e"], sc["q2"], sc["e2"] = sc["q2"], sc["e2"], sc["q"], sc["e"]
        mat = synth.score_matrix(sc["a"], sc["b"])
        w = int(rng.choice([-1, 5, 10, 20, 33, 37, 100, 150, 400]))
        flag = FLAGS[i % len(FLAGS)]
        args = (q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag)
        assert _same(ref_avx.ksw_extd2(*args), oracle.ksw_extd2(*args, score_rule=1)), "avx512 case %d" % i
        assert _same(ref_avx.ksw_extd2(*args, which=0), oracle.ksw_extd2(*args, score_rule=0)), "sse case %d" % i


def test_ksw_oracle_vs_reference_narrow_bands_random_scoring(oracle, ref_avx):
    """bands 0..5 with scorings that let a walk leave the band on the left (b > q + 2e): the AVX-512 build reads its
    lead-in cells there (off[r] rounded down to 64), the SSE build forces an insertion; both must be reproduced"""
    rng = np.random.default_rng(64)
    n_diff = 0
    for it in range(160):
        e, q = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        sc = dict(a=int(rng.integers(1, 5)), b=int(rng.integers(q + 2 * e + 1, 2 * (q + e) + 1)), q=q, e=e,
                  q2=int(rng.integers(6, 30)), e2=1, zdrop=int(rng.choice([400, 40])), end_bonus=int(rng.choice([0, 5])))
        mat = synth.score_matrix(sc["a"], sc["b"])
        P = synth.ragged_pairs(40, seed=int(rng.integers(1 << 30)), max_len=300)
        flag = int(rng.choice([0x08, 0x00, 0x0a, 0x88, 0x40, 0x18]))
        for i in range(P["n"]):
            qq, tt = pair(P, i)
            args = (qq, tt, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], int(rng.choice([0, 1, 2, 3, 5])), sc["zdrop"], sc["end_bonus"], flag)
            ra, rs = ref_avx.ksw_extd2(*args), ref_avx.ksw_extd2(*args, which=0)
            assert _same(ra, oracle.ksw_extd2(*args, score_rule=1)), "avx512 round %d case %d" % (it, i)
            assert _same(rs, oracle.ksw_extd2(*args, score_rule=0)), "sse round %d case %d" % (it, i)
            n_diff += 7 not in qq and not _same(ra, rs)
    assert n_diff >= 8  # the sweep does reach the lead-in cells


def test_ksw_oracle_vs_reference_long_band(oracle, ref_avx):
    P = synth.long_pairs(2, 3000, 0.08, seed=5, tlen_extra=0.02)
    sc = synth.SCORING["map-ont"]
    mat = synth.score_matrix(sc["a"], sc["b"])
    for i in range(P["n"]):
        q, t = pair(P, i)
        for flag, w in ((0x08, 200), (0x00, 150)):
            args = (q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag)
            assert _same(ref_avx.ksw_extd2(*args), oracle.ksw_extd2(*args))


def _seq(rng, n, kind):
    c = rng.integers(0, 4, n)
    if kind == 1:
        c[rng.random(n) < 0.01] = 4
    if kind == 2:
        for _ in range(3):
            a = int(rng.integers(0, n))
            c[a:a + int(rng.integers(1, 40))] = 4
    return bytes(synth.ACGTN[c])


def test_sketch_oracle_vs_reference(oracle, ref_avx, ref_scalar):
    """AVX-512 build (parity target) for w >= 8; scalar build additionally for mm_sketch2/3 (its
    mm_sketch differs by the documented '>' vs '>=' end rule, SURVEY.md finding 5)."""
    rng = np.random.default_rng(8)
    cfgs = [(21, 11), (19, 19), (15, 10), (28, 8), (17, 30), (11, 9), (12, 16)]
    pats = ["10", "110", "1110", "100", "11", "101001", "1"]
    for it in range(500):
        seq = _seq(rng, int(rng.choice([60, 150, 151, 300, 1000, 5000])), it % 3)
        k, w = cfgs[it % len(cfgs)]
        Z = pats[(it // 7) % len(pats)]
        assert np.array_equal(oracle.mm_sketch(seq, w, k, 3, Z), ref_avx.mm_sketch(seq, w, k, 3, Z)), it
        shift, cap = int(rng.integers(0, len(Z))), int(rng.choice([0, 3, 8, 800, 2 ** 32 - 1]))
        o3 = oracle.mm_sketch3(seq, w, k, 0, Z, shift, cap)
        for R in (ref_avx, ref_scalar):
            r3 = R.mm_sketch3(seq, w, k, 0, Z, shift, cap)
            assert np.array_equal(o3[0], r3[0]) and o3[1] == r3[1], (it, R.variant)
        ms = float(rng.choice([0.1, 0.2, 0.5, 1, 5, 50]))
        o2 = oracle.mm_sketch2(seq, w, k, 0, Z, ms)
        for R in (ref_avx, ref_scalar):
            r2 = R.mm_sketch2(seq, w, k, 0, Z, ms)
            assert np.array_equal(o2[0], r2[0]) and np.array_equal(o2[1], r2[1]), (it, R.variant)


def test_exact_match_vs_reference(oracle, ref_scalar):
    rng = np.random.default_rng(2)
    for n in (1, 15, 16, 17, 150, 299):
        a = rng.integers(0, 5, n).astype(np.uint8)
        assert oracle.exact_match(a, a.copy()) == ref_scalar.exact_match(a, a.copy()) == 1
        b = a.copy()
        b[int(rng.integers(0, n))] ^= 1
        assert oracle.exact_match(a, b) == ref_scalar.exact_match(a, b) == 0
