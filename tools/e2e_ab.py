#!/usr/bin/env python
"""A/B of the host-buffer DP call (gd_ksw_extd2_batch) with different pipeline slice sizes, same box, same data."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gdiet_b200 as gd
from gdiet_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ctx = gd.Context(0)
P = synth.ksw_pairs_fast(n, 150, 200, 0.05, seed=3)
sc = synth.SCORING["sr"]
prm = gd.KswParams(synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], sc["zdrop"], sc["end_bonus"], 0)
hp = {k: torch.from_numpy(P[k]).pin_memory() for k in ("qlen", "qoff", "qbuf", "tlen", "toff", "tbuf")}
out = {"ez": torch.zeros(n * 16, dtype=torch.int32).pin_memory().numpy().view(gd.GD_EXTZ_DTYPE),
       "cigar_off": torch.zeros(n + 1, dtype=torch.int64).pin_memory().numpy(),
       "cigar": torch.zeros(n * 24, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
def step():
    return ctx.ksw_extd2_batch(hp["qlen"].numpy(), hp["qoff"].numpy(), hp["qbuf"].numpy(), hp["tlen"].numpy(), hp["toff"].numpy(),
                               hp["tbuf"].numpy(), prm, w_all=150, out=out)
ref = None
for sl in (n, 125000, 0, 200000):
    ctx.set_option("ksw_slice", sl)
    for _ in range(2): step()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); ez, coff, cig = step(); ts.append((time.perf_counter() - t0) * 1e3)
    dt = min(ts) / 1e3
    print("   reps ms:", " ".join("%.1f" % x for x in ts))
    ck = (int(ez["score"].astype(np.int64).sum()), int(coff[-1]), int(cig[: int(coff[-1])].astype(np.int64).sum()))
    if ref is None: ref = ck
    print("slice %8d  %.2f ms  %.1f GCUPS  same=%s" % (sl, dt * 1e3, n * 28775 / dt / 1e9, ck == ref), flush=True)
