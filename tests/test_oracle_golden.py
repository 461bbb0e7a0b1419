"""CPU: the oracle restatement against the committed golden vectors (generated from the unmodified
reference's GDiet_avx objects by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oraclelib import EXTZ_FIELDS

import gdiet_b200  # noqa: F401
from gdiet_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCORINGS = ["sr", "map-hifi", "map-ont"]


def load_ksw_golden():
    return np.load(os.path.join(GOLD, "ksw_golden.npz"))


def load_sketch_golden():
    return np.load(os.path.join(GOLD, "sketch_golden.npz"))


def test_ksw_oracle_matches_golden(oracle):
    g = load_ksw_golden()
    n = len(g["qlen"])
    assert n >= 250
    for i in range(n):
        q = g["qbuf"][g["qoff"][i]:g["qoff"][i] + g["qlen"][i]]
        t = g["tbuf"][g["toff"][i]:g["toff"][i] + g["tlen"][i]]
        flag, sci, w = (int(x) for x in g["meta"][i])
        sc = synth.SCORING[SCORINGS[sci]]
        ez, cig = oracle.ksw_extd2(q, t, synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], w,
                                   sc["zdrop"], sc["end_bonus"], flag)
        assert [ez[f] for f in EXTZ_FIELDS] == [int(x) for x in g["ez"][i]], "case %d" % i
        assert np.array_equal(cig, g["cigar"][g["cigar_off"][i]:g["cigar_off"][i + 1]]), "case %d" % i


def test_ksw_oracle_matches_lead64_golden(oracle):
    """narrow bands + scorings with b > q + 2e: the walk reads the AVX-512 build's lead-in cells (off[r] rounded down to
    64, ksw2_extd2_avx.c:242,442); on half of the cases the SSE build returns a different CIGAR"""
    from helpers import lead64_golden_cases
    cases = lead64_golden_cases()
    assert sum(c[5] for c in cases) >= 40
    n16 = 0
    for i, (q, t, sc, flag, w, differ, ez_exp, cig_exp) in enumerate(cases):
        args = (q, t, synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag)
        ez, cig = oracle.ksw_extd2(*args)
        assert [ez[f] for f in EXTZ_FIELDS] == ez_exp and np.array_equal(cig, cig_exp), "case %d" % i
        ez16, cig16 = oracle.ksw_extd2(*args, score_rule=3)  # the 16-aligned rows of the SSE build
        n16 += [ez16[f] for f in EXTZ_FIELDS] != ez_exp or not np.array_equal(cig16, cig_exp)
    assert n16 == sum(c[5] for c in cases)


def test_sketch_oracle_matches_golden(oracle):
    g = load_sketch_golden()
    pats = [str(p) for p in g["patterns"]]
    n = len(g["meta"])
    assert n >= 200
    seq_all = g["seq"].tobytes()
    for i in range(n):
        kind, k, w, zi, fn, a0, a1 = (int(x) for x in g["meta"][i])
        seq = seq_all[g["seq_off"][i]:g["seq_off"][i + 1]]
        exp = g["entries"][g["entries_off"][i]:g["entries_off"][i + 1]]
        extra = g["extra"][g["extra_off"][i]:g["extra_off"][i + 1]]
        Z = pats[zi]
        if fn == 0:
            got = oracle.mm_sketch(seq, w, k, a0, Z)
        elif fn == 1:
            got, ret = oracle.mm_sketch3(seq, w, k, 0, Z, a0, a1)
            assert ret == int(extra[0]), "case %d ret" % i
        else:
            got, counts = oracle.mm_sketch2(seq, w, k, 0, Z, a0 / 1000.0)
            assert np.array_equal(counts, extra), "case %d counts" % i
        assert np.array_equal(got, exp), "case %d (fn %d kind %d k %d w %d Z %s)" % (i, fn, kind, k, w, Z)


def test_band_cells_closed_forms(oracle):
    # full band = qlen*tlen; 150x200, w=150 = 28,775
    assert oracle.band_cells(150, 150, 150) == 22500
    assert oracle.band_cells(150, 200, 150) == 28775
    assert oracle.band_cells(10, 10, -1) == 100


def test_exact_match(oracle):
    a = np.array([0, 1, 2, 3, 4, 0, 1], np.uint8)
    assert oracle.exact_match(a, a.copy()) == 1
    b = a.copy()
    b[6] = 2
    assert oracle.exact_match(a, b) == 0


def test_map_oracle_matches_golden():
    """oracle/gd_oracle_map.c against tests/golden/map_sr.npz (generated from the reference program's call trace by
    tests/golden/make_golden_map.py): works on machines without oracle/_ref."""
    import maplib
    g = np.load(os.path.join(GOLD, "map_sr.npz"))
    contigs, reads = maplib.make_dataset(seed=int(g["seed"]), n_reads=int(g["n_reads"]))
    o = maplib.sr_opt(min_cnt=float(g["min_cnt"]), rec_frac=float(g["rec_frac"]))
    M = maplib.MapOracle()
    mi = M.index_build(contigs, 11, 21, "10")
    k = 0
    cig_at = 0
    for i, r in enumerate(reads):
        c, cig, _ = M.map_read(mi, r, o)
        assert len(c) == g["cand_off"][i + 1] - g["cand_off"][i], "read %d" % i
        for a in c:
            for f in ("rid", "rs", "re", "qs", "qe", "rev", "exact", "score", "n_cigar"):
                assert int(a[f]) == int(g[f][k]), "read %d field %s" % (i, f)
            n = max(int(a["n_cigar"]), 0)
            assert np.array_equal(cig[int(a["cigar_off"]):int(a["cigar_off"]) + n], g["cigar"][cig_at:cig_at + n])
            cig_at += n
            k += 1
    assert k == len(g["rid"]) and k > 1500
    M.lib.gdo_index_destroy(mi)
