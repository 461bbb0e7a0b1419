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
