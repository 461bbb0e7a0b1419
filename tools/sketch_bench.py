#!/usr/bin/env python
"""Developer timing of the sparsified-sketch kernel (BASELINE config 5 shapes): index sketching of a synthetic
genome (mm_sketch, -Z 10 -W 2 -k 21 -w 11), device resident, and read sketching (mm_sketch2 + mm_sketch3, all
shifts) of 150 bp reads through the host-buffer ABI.  Prints Gbases/s and the kernel's HBM roofline fraction
(algorithmic bytes = 1 B per input base + 16 B per emitted minimizer)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gdiet_b200 as gd
from gdiet_b200 import synth

def main():
    mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    nreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    ctx = gd.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # ---- index sketching: ncontig contigs of 25 Mbp
    L = 25_000_000
    ncontig = max(1, mbp // 25)
    g = torch.randint(0, 4, (ncontig * L,), dtype=torch.uint8, device=dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    seq = lut[g.long()] if ncontig * L < 2**31 else None
    d_off = torch.arange(ncontig, dtype=torch.int64, device=dev) * L
    d_len = torch.full((ncontig,), L, dtype=torch.int32, device=dev)
    d_rid = torch.arange(ncontig, dtype=torch.int32, device=dev)
    cap = ncontig * L // 5
    d_out = torch.zeros(cap * 2, dtype=torch.int64, device=dev)
    d_oo = torch.zeros(ncontig + 1, dtype=torch.int64, device=dev)
    ctx.set_option("time_kernels", 1)
    for it in range(4):
        ctx.stat("sketch_reset")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.sketch_ref_batch_device(ncontig, d_off, d_len, d_rid, seq, ncontig * L, 11, 21, "10", d_oo, d_out, cap)
        stream.synchronize()
        dt = time.perf_counter() - t0
        us = ctx.stat("sketch_us")
    nmin = int(d_oo[-1].item())
    bases = ncontig * L
    algo = bases + 16 * nmin
    print(json.dumps({"what": "index sketch (mm_sketch), device resident", "bases": bases, "minimizers": nmin,
                      "gbases_s_call": bases / dt / 1e9, "kernel_ms": us / 1e3, "gbases_s_kernel": bases / (us * 1e-6) / 1e9,
                      "algo_gbs_kernel": algo / (us * 1e-6) / 1e9, "hbm_frac": algo / (us * 1e-6) / 1e9 / hbm}), flush=True)
    del g, seq, d_out
    torch.cuda.empty_cache()
    # ---- read sketching through the host-buffer ABI
    genome = synth.random_genome(5_000_000, seed=1)
    reads = synth.sample_reads(genome, nreads, 150, seed=8)
    off = np.arange(nreads, dtype=np.int64) * 150
    lens = np.full(nreads, 150, np.int32)
    buf = np.ascontiguousarray(reads.reshape(-1))
    for it in range(3):
        ctx.stat("sketch_reset")
        t0 = time.perf_counter()
        R = ctx.sketch_reads_batch(off, lens, buf, 11, 21, "10", 0.1, 800)
        dt = time.perf_counter() - t0
        us = ctx.stat("sketch_us")
    print(json.dumps({"what": "read sketch (mm_sketch2 + mm_sketch3 for both shifts), host buffers", "reads": nreads, "bases": nreads * 150,
                      "s2": int(R["s2_off"][-1]), "s3": int(R["s3_off"][-1]), "gbases_s_call": nreads * 150 / dt / 1e9,
                      "kernel_ms": us / 1e3, "gbases_s_kernel": nreads * 150 / (us * 1e-6) / 1e9}), flush=True)

if __name__ == "__main__":
    main()
