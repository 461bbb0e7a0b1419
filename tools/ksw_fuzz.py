"""Randomised parity sweep of the CUDA DP against the oracle: scoring values, flags, bands, lengths and gang sizes
drawn at random per round (the parity tests in tests/test_gpu_ksw.py fix these to the reference's presets).
Test infrastructure: the oracle is the checker only.  Prints one JSON line per round and a summary line.

  python tools/ksw_fuzz.py --seconds 150 --seed 1 > gpurun_out/fuzz.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gdiet_b200 as gd  # noqa: E402
from gdiet_b200 import synth  # noqa: E402
from helpers import oracle_batch, params  # noqa: E402
from oraclelib import EXTZ_FIELDS, Oracle  # noqa: E402

FLAGS = [0x00, 0x01, 0x02, 0x08, 0x0a, 0x10, 0x18, 0x40, 0x42, 0x48, 0x80, 0x88, 0xc2, 0xc8, 0x58]


def draw_scoring(rng):
    """values inside the int8 difference form's range (ksw2_extd2_sse.c:54-58: q+e, q2+e2 and the match score
    are stored as int8 lanes), both piece orders, zero and large Z-drop / end bonus"""
    a = int(rng.integers(1, 7))
    b = int(rng.integers(1, 13))
    q = int(rng.integers(1, 30))
    e = int(rng.integers(1, 6))
    q2 = int(rng.integers(1, 60))
    e2 = int(rng.integers(1, 4))
    zdrop = int(rng.choice([-1, 0, 5, 40, 100, 400, 2000]))
    end_bonus = int(rng.choice([-1, 0, 1, 10, 50]))
    return dict(a=a, b=b, q=q, e=e, q2=q2, e2=e2, zdrop=zdrop, end_bonus=end_bonus)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--pairs", type=int, default=200)
    ap.add_argument("--narrow", action="store_true", help="bands 0..5 and scorings with b > q + 2e: walks that read the AVX-512 build's lead-in cells")
    ap.add_argument("--long", action="store_true", help="long pairs (1..12 kbp, bands 50..1300): block-per-pair gangs, ring wrap")
    a = ap.parse_args()
    O = Oracle()
    ctx = gd.Context(0)
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    rounds = pairs = bad = redone = 0
    while time.time() - t0 < a.seconds:
        sc = draw_scoring(rng)
        flag = int(rng.choice(FLAGS))
        max_len = int(rng.choice([8, 40, 150, 300, 700]))
        if a.long:
            max_len = int(rng.choice([1000, 3000, 6000, 12000]))
            P = synth.long_pairs(int(rng.integers(2, 9)), max_len, float(rng.choice([0.01, 0.08, 0.15])),
                                 seed=int(rng.integers(1 << 30)), tlen_extra=float(rng.choice([0.0, 0.01, 0.05])))
            w = rng.choice([50, 151, 500, 1000, 1300], P["n"]).astype(np.int32)
        else:
            P = synth.ragged_pairs(a.pairs, seed=int(rng.integers(1 << 30)), max_len=max_len)
            w = rng.choice([-1, 0, 1, 3, 5, 10, 20, 33, 37, 64, 100, 150, 400, 1000], P["n"]).astype(np.int32)
        if a.narrow:
            e, q = int(rng.integers(1, 3)), int(rng.integers(1, 4))
            sc.update(b=int(rng.integers(q + 2 * e + 1, 2 * (q + e) + 1)), q=q, e=e, q2=int(rng.integers(6, 30)), e2=1)
            w = rng.choice([0, 1, 2, 3, 5], P["n"]).astype(np.int32)
        G = int(rng.choice([0, 4, 8, 16, 32]))
        exp = oracle_batch(O, P, w, sc, flag)
        ctx.set_option("ksw_group", G)
        ez, coff, cig = ctx.ksw_extd2_batch(P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                            params(sc, flag), w=w)
        mism = []
        for i, (eo, co) in enumerate(exp):
            mine = {f: int(ez[i][f]) for f in EXTZ_FIELDS}
            ok = mine == eo
            if ok and not (flag & 1):
                ok = np.array_equal(cig[int(coff[i]):int(coff[i + 1])], co)
            if not ok:
                mism.append(dict(pair=i, qlen=int(P["qlen"][i]), tlen=int(P["tlen"][i]), w=int(w[i]), got=mine, exp=eo))
        rounds += 1
        pairs += P["n"]
        bad += len(mism)
        redone += int(ez["lead64"].sum())
        print(json.dumps(dict(round=rounds, scoring=sc, flag=flag, max_len=max_len, G=G, pairs=P["n"],
                              mismatches=len(mism), first=mism[:2])), flush=True)
    print(json.dumps(dict(summary=True, seed=a.seed, rounds=rounds, pairs=pairs, mismatches=bad, lead64_redone=redone,
                          seconds=round(time.time() - t0, 1))), flush=True)
    ctx.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
