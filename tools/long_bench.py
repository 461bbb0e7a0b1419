#!/usr/bin/env python
"""Developer timing of the long-read DP shapes of BASELINE configs 3/4 (HiFi 15 kbp w=1000, ONT 50 kbp w=1300).
Prints banded GCUPS of the device-resident batched call (pack + DP + traceback) and of the DP kernel alone."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gdiet_b200 as gd
from gdiet_b200 import synth

def cells(qlen, tlen, w):
    r = np.arange(qlen + tlen - 1)
    st0 = np.maximum(np.maximum(0, r - qlen + 1), (r - w + 1) >> 1)
    en0 = np.minimum(np.minimum(tlen - 1, r), (r + w) >> 1)
    return int(np.maximum(en0 - st0 + 1, 0).sum())

GROUP = int(os.environ.get("GD_GROUP", "0"))


def main():
    ctx = gd.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for name, n, qlen, edit, w, scn, flag in (("hifi", 2048, 15000, 0.01, 1000, "map-hifi", 0x08), ("ont", 512, 50000, 0.08, 1300, "map-ont", 0x08),
                                              ("hifi-exact", 2048, 15000, 0.01, 1000, "map-hifi", 0x00)):
        P = synth.long_pairs(8, qlen, edit, seed=7, tlen_extra=0.01)
        # replicate the 8 generated pairs to n (timing only)
        rep = n // 8
        qbuf, tbuf = np.tile(P["qbuf"], rep), np.tile(P["tbuf"], rep)
        ql, tl = np.tile(P["qlen"], rep), np.tile(P["tlen"], rep)
        qoff = np.concatenate([[0], np.cumsum(ql[:-1])]).astype(np.int64)
        toff = np.concatenate([[0], np.cumsum(tl[:-1])]).astype(np.int64)
        sc = synth.SCORING[scn]
        prm = gd.KswParams(synth.score_matrix(sc["a"], sc["b"]), sc["q"], sc["e"], sc["q2"], sc["e2"], sc["zdrop"], sc["end_bonus"], flag)
        d = {k: torch.from_numpy(v).to(dev) for k, v in dict(qlen=ql, qoff=qoff, qbuf=qbuf, tlen=tl, toff=toff, tbuf=tbuf).items()}
        stride = int(ql.max() + tl.max())
        d_ez = torch.zeros(n * 16, dtype=torch.int32, device=dev)
        d_cig = torch.zeros(n * 4096, dtype=torch.int32, device=dev)
        ctx.set_option("time_kernels", 1)
        ctx.set_option("ksw_group", GROUP)
        tot = sum(cells(int(a), int(b), w) for a, b in zip(ql, tl))
        for it in range(3):
            ctx.stat("ksw_dp_reset")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.ksw_extd2_batch_device(n, d["qlen"], d["qoff"], d["qbuf"], d["tlen"], d["toff"], d["tbuf"], prm, int(ql.max()), int(tl.max()), w,
                                       d_ez, d_cig, 4096, w_all=w)
            stream.synchronize()
            dt = time.perf_counter() - t0
            us = ctx.stat("ksw_dp_us")
        ez = d_ez.cpu().numpy().view(gd.GD_EXTZ_DTYPE)
        print(json.dumps({"shape": name, "pairs": n, "qlen": qlen, "w": w, "flag": flag, "step_ms": dt * 1e3, "dp_ms": us / 1e3,
                          "gcups_step": tot / dt / 1e9, "gcups_dp": tot / (us * 1e-6) / 1e9, "group": ctx.stat("ksw_group"),
                          "ring": ctx.stat("ksw_ring"), "chunks": ctx.stat("ksw_chunks"), "score0": int(ez["score"][0]),
                          "zdropped": int((ez["zdropped"] != 0).sum())}), flush=True)

if __name__ == "__main__":
    main()
