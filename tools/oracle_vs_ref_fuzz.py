"""Randomised check of the oracle's DP restatement against the compiled reference (oracle/_ref, AVX-512 build) with
scoring values, flags and bands the committed fixtures do not hold.  CPU only; needs oracle/_ref (built in the
container that has /root/reference).  Test infrastructure: compares the two checkers, touches no product code.

  python tools/oracle_vs_ref_fuzz.py --seconds 60 [--narrow]     # --narrow: bands 0..3 only
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

from gdiet_b200 import synth  # noqa: E402
from helpers import oracle_batch  # noqa: E402
from oraclelib import Oracle, Ref, cpu_has_avx512  # noqa: E402
import ksw_fuzz as kf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--narrow", action="store_true")
    a = ap.parse_args()
    O, R = Oracle(), Ref("avx" if cpu_has_avx512() else "scalar")
    rng = np.random.default_rng(a.seed)
    bands = [0, 1, 2, 3] if a.narrow else [-1, 0, 1, 3, 5, 10, 20, 33, 37, 64, 100, 150, 400, 1000]
    n = {"preset": 0, "random": 0}
    bad = {"preset": 0, "random": 0}
    t0 = time.time()
    while time.time() - t0 < a.seconds:
        kind = "preset" if rng.random() < 0.4 else "random"
        sc = dict(synth.SCORING[str(rng.choice(["sr", "map-hifi", "map-ont"]))]) if kind == "preset" else kf.draw_scoring(rng)
        flag = int(rng.choice(kf.FLAGS))
        P = synth.ragged_pairs(60, seed=int(rng.integers(1 << 30)), max_len=int(rng.choice([8, 40, 150, 300])))
        w = rng.choice(bands, P["n"]).astype(np.int32)
        x, y = oracle_batch(O, P, w, sc, flag), oracle_batch(R, P, w, sc, flag)
        for i, (p, q) in enumerate(zip(x, y)):
            n[kind] += 1
            if p[0] != q[0] or not np.array_equal(p[1], q[1]):
                bad[kind] += 1
                print(json.dumps(dict(kind=kind, scoring=sc, flag=flag, w=int(w[i]), qlen=int(P["qlen"][i]),
                                      tlen=int(P["tlen"][i]), ez_equal=p[0] == q[0])), flush=True)
    print(json.dumps(dict(summary=True, pairs=n, differ=bad, narrow=a.narrow, seed=a.seed)))


if __name__ == "__main__":
    main()
