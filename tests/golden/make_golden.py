"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref/libgdref_avx.so, i.e. the objects
of GDiet_avx: ksw_extd2_avx512 + AVX-512 mm_sketch*).  Run in the build container:

    python tests/golden/make_golden.py

Writes tests/golden/ksw_golden.npz and tests/golden/sketch_golden.npz (committed). The reference ships
no tests of its own, so these files pin parity for machines where /root/reference and
oracle/_ref do not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oraclelib import Ref, EXTZ_FIELDS  # noqa: E402
import gdiet_b200  # noqa: E402,F401
from gdiet_b200 import synth  # noqa: E402

FLAGS = [0x08, 0x00, 0x18, 0x40, 0x48, 0xc2, 0x01, 0x09, 0x0a, 0x42, 0x80]
SCORINGS = ["sr", "map-hifi", "map-ont"]
BANDS = [-1, 5, 10, 20, 33, 37, 100, 150, 400]


def ksw_cases():
    rng = np.random.default_rng(20261018)
    R = Ref("avx")
    groups = [synth.ragged_pairs(160, seed=101, max_len=260), synth.ksw_pairs(48, 150, 200, 0.05, seed=3),
              synth.ksw_pairs(16, 150, 200, 0.30, seed=4), synth.ksw_pairs(24, 150, 150, 0.02, seed=5),
              synth.long_pairs(6, 1200, 0.08, seed=6, tlen_extra=0.03)]
    # degenerate shapes
    groups.append(synth.pack_pairs([np.array([1], np.uint8), np.array([0, 1, 2, 3] * 5, np.uint8), np.array([2], np.uint8)],
                                   [np.array([1], np.uint8), np.array([3], np.uint8), np.array([0, 1, 2, 3] * 9, np.uint8)]))
    qs, ts, meta, ezs, cigs = [], [], [], [], []
    for gi, P in enumerate(groups):
        for i in range(P["n"]):
            q = P["qbuf"][P["qoff"][i]:P["qoff"][i] + P["qlen"][i]]
            t = P["tbuf"][P["toff"][i]:P["toff"][i] + P["tlen"][i]]
            flag = FLAGS[(i + gi) % len(FLAGS)]
            scn = SCORINGS[(i // 3 + gi) % 3]
            w = int(rng.choice(BANDS)) if gi in (0, 5) else (150 if gi in (1, 2, 3) else 100)
            sc = synth.SCORING[scn]
            mat = synth.score_matrix(sc["a"], sc["b"])
            ez, cig = R.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag)
            ez_sse, cig_sse = R.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"],
                                          flag, which=0)
            if 7 not in q:  # for codes 0..4 the SSE and AVX-512 kernels must agree (SURVEY.md finding 3)
                assert ez == ez_sse and np.array_equal(cig, cig_sse), (gi, i)
            qs.append(q), ts.append(t), cigs.append(cig)
            meta.append([flag, SCORINGS.index(scn), w])
            ezs.append([ez[f] for f in EXTZ_FIELDS])
    P = synth.pack_pairs(qs, ts)
    coff = np.zeros(len(cigs) + 1, np.int64)
    coff[1:] = np.cumsum([len(c) for c in cigs])
    np.savez_compressed(os.path.join(HERE, "ksw_golden.npz"), qbuf=P["qbuf"], qoff=P["qoff"], qlen=P["qlen"],
                        tbuf=P["tbuf"], toff=P["toff"], tlen=P["tlen"], meta=np.array(meta, np.int32),
                        ez=np.array(ezs, np.int32), cigar_off=coff,
                        cigar=np.concatenate(cigs).astype(np.uint32) if cigs else np.zeros(0, np.uint32))
    print("ksw golden:", len(qs), "cases")


def sketch_cases():
    rng = np.random.default_rng(77)
    R = Ref("avx")
    cfgs = [(21, 11), (19, 19), (15, 10), (28, 8), (17, 30)]
    pats = ["10", "110", "1110", "100", "11", "101001"]
    seqs, meta, outs = [], [], []

    def add(seq, kind, k, w, Z, fn, arg0, arg1, entries, extra):
        seqs.append(seq)
        meta.append([kind, k, w, pats.index(Z), fn, arg0, arg1])
        outs.append((entries, extra))

    for it in range(240):
        kind = it % 3  # 0 random, 1 with single Ns, 2 N runs
        n = int(rng.choice([60, 150, 151, 300, 1000, 6000]))
        codes = rng.integers(0, 4, n)
        if kind == 1:
            codes[rng.random(n) < 0.01] = 4
        if kind == 2:
            for _ in range(3):
                a = int(rng.integers(0, n))
                codes[a:a + int(rng.integers(1, 40))] = 4
        seq = bytes(synth.ACGTN[codes])
        k, w = cfgs[it % len(cfgs)]
        Z = pats[(it // 5) % len(pats)]
        fn = it % 3
        if fn == 0:
            e = R.mm_sketch(seq, w, k, it % 7, Z)
            add(seq, kind, k, w, Z, 0, it % 7, 0, e, np.zeros(0, np.uint32))
        elif fn == 1:
            shift = int(rng.integers(0, len(Z)))
            cap = int(rng.choice([0, 3, 8, 800, 2 ** 32 - 1]))
            e, ret = R.mm_sketch3(seq, w, k, 0, Z, shift, cap)
            add(seq, kind, k, w, Z, 1, shift, cap, e, np.array([ret], np.uint32))
        else:
            ms = float(rng.choice([0.1, 0.2, 0.5, 1, 5, 50]))
            e, counts = R.mm_sketch2(seq, w, k, 0, Z, ms)
            add(seq, kind, k, w, Z, 2, int(ms * 1000), 0, e, counts)
    soff = np.zeros(len(seqs) + 1, np.int64)
    soff[1:] = np.cumsum([len(s) for s in seqs])
    eoff = np.zeros(len(outs) + 1, np.int64)
    eoff[1:] = np.cumsum([len(o[0]) for o in outs])
    xoff = np.zeros(len(outs) + 1, np.int64)
    xoff[1:] = np.cumsum([len(o[1]) for o in outs])
    np.savez_compressed(os.path.join(HERE, "sketch_golden.npz"), seq=np.frombuffer(b"".join(seqs), np.uint8), seq_off=soff,
                        meta=np.array(meta, np.int64), entries=np.concatenate([o[0] for o in outs]).astype(np.uint64),
                        entries_off=eoff, extra=np.concatenate([o[1] for o in outs]).astype(np.uint32), extra_off=xoff,
                        patterns=np.array(pats))
    print("sketch golden:", len(seqs), "cases")




def ksw_lead64_cases():
    """Narrow bands + scorings with b > q + 2e: the walks that step off the band's left edge into the AVX-512 build's
    lead-in cells (off[r] rounded down to 64, ksw2_extd2_avx.c:242,442; read by ksw_backtrack, ksw2.h:136).  Kept: every
    pair on which ksw_extd2_avx512 and ksw_extd2_sse return different CIGARs, plus as many on which they agree.
    Every case carries its own scoring values.  Writes tests/golden/ksw_lead64_golden.npz."""
    rng = np.random.default_rng(64)
    R = Ref("avx")
    flags = [0x08, 0x00, 0x0a, 0x88, 0x40, 0x18]
    qs, ts, meta, ezs, cigs = [], [], [], [], []
    n_diff = n_same = 0
    while n_diff < 48:
        e = int(rng.integers(1, 3))
        q = int(rng.integers(1, 4))
        b = int(rng.integers(q + 2 * e + 1, 2 * (q + e) + 1))
        sc = dict(a=int(rng.integers(1, 5)), b=b, q=q, e=e, q2=int(rng.integers(6, 30)), e2=1, zdrop=int(rng.choice([400, 40])),
                  end_bonus=int(rng.choice([0, 5])))
        mat = synth.score_matrix(sc["a"], sc["b"])
        P = synth.ragged_pairs(40, seed=int(rng.integers(1 << 30)), max_len=300)
        flag = int(rng.choice(flags))
        for i in range(P["n"]):
            qq = P["qbuf"][P["qoff"][i]:P["qoff"][i] + P["qlen"][i]]
            tt = P["tbuf"][P["toff"][i]:P["toff"][i] + P["tlen"][i]]
            if 7 in qq:
                continue
            w = int(rng.choice([0, 1, 2, 3, 5]))
            ez, cig = R.ksw_extd2(qq, tt, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag)
            ez0, cig0 = R.ksw_extd2(qq, tt, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], w, sc["zdrop"], sc["end_bonus"], flag, which=0)
            differ = ez != ez0 or not np.array_equal(cig, cig0)
            if not differ and n_same >= n_diff:
                continue
            n_diff += differ
            n_same += not differ
            qs.append(qq), ts.append(tt), cigs.append(cig)
            meta.append([flag, w, sc["a"], sc["b"], sc["q"], sc["e"], sc["q2"], sc["e2"], sc["zdrop"], sc["end_bonus"], int(differ)])
            ezs.append([ez[f] for f in EXTZ_FIELDS])
    P = synth.pack_pairs(qs, ts)
    coff = np.zeros(len(cigs) + 1, np.int64)
    coff[1:] = np.cumsum([len(c) for c in cigs])
    np.savez_compressed(os.path.join(HERE, "ksw_lead64_golden.npz"), qbuf=P["qbuf"], qoff=P["qoff"], qlen=P["qlen"],
                        tbuf=P["tbuf"], toff=P["toff"], tlen=P["tlen"], meta=np.array(meta, np.int32),
                        ez=np.array(ezs, np.int32), cigar_off=coff, cigar=np.concatenate(cigs).astype(np.uint32))
    print("ksw lead64 golden:", len(qs), "cases,", n_diff, "with AVX-512 != SSE")


if __name__ == "__main__":
    if "lead64" in sys.argv[1:]:  # only the file added in round 2
        ksw_lead64_cases()
    else:
        ksw_cases()
        sketch_cases()
        ksw_lead64_cases()
