"""Randomised parity sweep of the CUDA sketching against the oracle: k, w, pattern, lengths, N runs, lower case,
seed caps and max_seeds drawn at random per round.  Test infrastructure: the oracle is the checker only.

  python tools/sketch_fuzz.py --seconds 100 --seed 1 > gpurun_out/sketch_fuzz.jsonl
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
from oraclelib import Oracle  # noqa: E402

ALPHA = np.frombuffer(b"ACGTNacgtnRY", np.uint8)


def _xy(a):
    return np.stack([a["x"], a["y"]], 1) if len(a) else np.zeros((0, 2), np.uint64)


def draw_seqs(rng, n, max_len):
    seqs = []
    for _ in range(n):
        ln = int(rng.integers(1, max_len + 1))  # both batch entries reject empty sequences (explicit error)
        c = rng.integers(0, 4, ln)
        mode = rng.random()
        if mode < 0.3 and ln:  # N runs
            for _ in range(int(rng.integers(1, 4))):
                a = int(rng.integers(0, ln))
                c[a:a + int(rng.integers(1, 40))] = 4
        elif mode < 0.4 and ln:  # lower case / IUPAC sprinkled
            m = rng.random(ln) < 0.05
            c = np.where(m, rng.integers(4, len(ALPHA), ln), c)
        elif mode < 0.5 and ln:  # low complexity: fw == rv palindromes and ties
            c = np.tile(rng.integers(0, 4, int(rng.integers(1, 5))), ln)[:ln]
        seqs.append(ALPHA[c].tobytes())
    return seqs


def pack(seqs):
    ln = np.array([len(s) for s in seqs], np.int32)
    off = np.zeros(len(seqs), np.int64)
    off[1:] = np.cumsum(ln[:-1])
    buf = np.frombuffer(b"".join(seqs) + b"A", np.uint8)[:-1] if sum(ln) else np.zeros(0, np.uint8)
    return off, ln, np.ascontiguousarray(buf)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=100.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    O = Oracle()
    ctx = gd.Context(0)
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    rounds = seqs_n = bad = 0
    while time.time() - t0 < a.seconds:
        k = int(rng.integers(1, 29))
        w = int(rng.choice([1, 2, 5, 10, 11, 19, 30, 64, 200, 255]))
        W = int(rng.integers(1, 9))
        Z = "".join(rng.choice(["0", "1"], W))
        if "1" not in Z:
            Z = "1" + Z[1:]
        reads_mode = bool(rng.random() < 0.5) and W > 1  # (the reads entry refuses W = 1: mm_sketch2 is not pinned for it)
        max_len = int(rng.choice([30, 150, 400, 3000] if reads_mode else [30, 400, 3000, 20000]))
        seqs = [s for s in draw_seqs(rng, 24, max_len)]
        if reads_mode:
            seqs = [s if len(s) >= W else s + b"ACGTACGT"[:W - len(s)] for s in seqs]  # the reads entry rejects reads shorter than the pattern
        off, ln, buf = pack(seqs)
        mism = []
        info = dict(k=k, w=w, Z=Z, max_len=max_len)
        if not reads_mode:
            rid = rng.integers(0, 1 << 20, len(seqs)).astype(np.uint32)
            oo, out = ctx.sketch_ref_batch(off, ln, buf, w, k, Z, rid=rid)
            for i, s in enumerate(seqs):
                if not np.array_equal(_xy(out[oo[i]:oo[i + 1]]), O.mm_sketch(s, w, k, int(rid[i]), Z)):
                    mism.append(dict(fn="mm_sketch", i=i, len=len(s)))
        else:
            ms = float(rng.choice([0.1, 0.5, 1.0, 2.0, 5.0]))
            cap = int(rng.choice([1, 3, 40, 800, 0xffffffff]))
            info.update(max_seeds=ms, max_nb_seeds=cap)
            R = ctx.sketch_reads_batch(off, ln, buf, w, k, Z, ms, cap)
            for i, s in enumerate(seqs):
                e2, c2 = O.mm_sketch2(s, w, k, 0, Z, ms)
                if not (np.array_equal(R["s2_counts"][i], c2)
                        and np.array_equal(_xy(R["s2"][R["s2_off"][i]:R["s2_off"][i + 1]]), e2)):
                    mism.append(dict(fn="mm_sketch2", i=i, len=len(s)))
                for sh in range(W):
                    e3, ret = O.mm_sketch3(s, w, k, 0, Z, sh, cap)
                    q = i * W + sh
                    if not (np.array_equal(_xy(R["s3"][R["s3_off"][q]:R["s3_off"][q + 1]]), e3)
                            and int(R["s3_ret"][i, sh]) == ret):
                        mism.append(dict(fn="mm_sketch3", i=i, len=len(s), shift=sh))
        rounds += 1
        seqs_n += len(seqs)
        bad += len(mism)
        print(json.dumps(dict(round=rounds, reads=reads_mode, mismatches=len(mism), first=mism[:3], **info)), flush=True)
    print(json.dumps(dict(summary=True, seed=a.seed, rounds=rounds, sequences=seqs_n, mismatches=bad,
                          seconds=round(time.time() - t0, 1))), flush=True)
    ctx.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
