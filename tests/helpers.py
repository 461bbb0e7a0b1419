"""Shared helpers for the parity tests."""
import numpy as np

from oraclelib import EXTZ_FIELDS

import gdiet_b200 as gd
from gdiet_b200 import synth


def pair(P, i):
    q = P["qbuf"][P["qoff"][i]:P["qoff"][i] + P["qlen"][i]]
    t = P["tbuf"][P["toff"][i]:P["toff"][i] + P["tlen"][i]]
    return q, t


def oracle_batch(O, P, w, sc, flag, idx=None):
    """Run the oracle over (a subset of) a batch; returns list of (ez dict, cigar)."""
    mat = synth.score_matrix(sc["a"], sc["b"])
    out = []
    for i in (range(P["n"]) if idx is None else idx):
        q, t = pair(P, i)
        out.append(O.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], int(w[i]), sc["zdrop"], sc["end_bonus"],
                               flag))
    return out


def assert_batch_equal(ez, coff, cig, expected, idx=None, what=""):
    idx = list(range(len(expected))) if idx is None else list(idx)
    for (eo, co), i in zip(expected, idx):
        mine = {f: int(ez[i][f]) for f in EXTZ_FIELDS}
        assert mine == eo, "%s pair %d ez mismatch:\n got %s\n exp %s" % (what, i, mine, eo)
        if cig is not None:
            got = cig[int(coff[i]):int(coff[i + 1])]
            assert np.array_equal(got, co), "%s pair %d cigar mismatch:\n got %s\n exp %s" % (what, i, got, co)


def params(sc, flag):
    return gd.KswParams(synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], sc["zdrop"],
                        sc["end_bonus"], flag)


def cigar_spans(cig):
    """(query span, target span) consumed by a BAM cigar array."""
    ops = cig & 0xf
    lens = (cig >> 4).astype(np.int64)
    return int(lens[(ops == 0) | (ops == 1)].sum()), int(lens[(ops == 0) | (ops == 2)].sum())


def lead64_golden_cases():
    """tests/golden/ksw_lead64_golden.npz -> (q, t, scoring dict, flag, w, differs-from-SSE, ez list, cigar) per case"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ksw_lead64_golden.npz"))
    out = []
    for i in range(len(g["qlen"])):
        q = g["qbuf"][g["qoff"][i]:g["qoff"][i] + g["qlen"][i]]
        t = g["tbuf"][g["toff"][i]:g["toff"][i] + g["tlen"][i]]
        flag, w, a, b, gq, ge, gq2, ge2, zdrop, end_bonus, differ = (int(x) for x in g["meta"][i])
        sc = dict(a=a, b=b, q=gq, e=ge, q2=gq2, e2=ge2, zdrop=zdrop, end_bonus=end_bonus)
        out.append((q, t, sc, flag, w, differ, [int(x) for x in g["ez"][i]],
                    g["cigar"][g["cigar_off"][i]:g["cigar_off"][i + 1]]))
    return out
