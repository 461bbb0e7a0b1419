#!/usr/bin/env python
"""BASELINE config 1 end to end: `-ax sr -Z 10 -W 2 -k 21 -w 11 -r 0.05,150,200`, synthetic 150 bp reads against a
synthetic reference.  Ours: index build on the device, then per batch gd_sr_map_batch (sketch -> lookup -> vote ->
windows -> exact match / DP, all on the GPU) + gd_sr_sam_batch (post-processing + SAM text on the host cores); the
SAM text is compared with the unmodified reference program (oracle/_ref/GDiet_avx_sr -t <cores>) run on the same
files in the same process, whose [PROFILING]/Real-time lines give the CPU baseline.  Prints one JSON line."""
import json, os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gdiet_b200 as gd
from gdiet_b200 import synth


def run(ctx, ref_mbp=5, n_reads=100_000, run_ref=True):
    cores = len(os.sched_getaffinity(0))
    genome = synth.random_genome(int(ref_mbp * 1e6), seed=1)
    reads = synth.sample_reads(genome, n_reads, 150, seed=2)
    contigs = [genome]
    name_list = ["r%d" % i for i in range(n_reads)]
    names = gd._cstr_array(name_list)  # a C host already holds char* names
    nb_ = max(1, min(8, n_reads // 125_000))
    C_names = [gd._cstr_array(name_list[n_reads * b // nb_: n_reads * (b + 1) // nb_]) for b in range(nb_)]
    off = np.arange(n_reads, dtype=np.int64) * 150
    lens = np.full(n_reads, 150, np.int32)
    buf = np.ascontiguousarray(reads.reshape(-1))
    qual = np.full(n_reads * 150, ord("I"), np.uint8)
    # ---- whole programs, file to file, BEFORE this process takes device memory for its own contexts (a second process next to a
    # loaded one gets what is left of the HBM and shares the page-locking path): the unmodified reference, then the batched C host
    file_runs, want = {}, None
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "GDiet_avx_sr")
    batched_bin = os.path.join(ROOT, "oracle", "_ref", "GDiet_cuda_batched_sr")
    if run_ref and os.path.exists(ref_bin):
        import maplib
        tmp = tempfile.mkdtemp(prefix="gdref_")
        fa, fq, samf = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq"), os.path.join(tmp, "out.sam")
        maplib.write_fasta(fa, contigs)
        maplib.write_fastq(fq, reads)
        t0 = time.perf_counter()
        p = subprocess.run([ref_bin, "-t", str(cores), "-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.05,150,200",
                            "-o", samf, fa, fq], capture_output=True, text=True)
        wall = time.perf_counter() - t0
        prof = dict(re.findall(r"\[PROFILING\] (.+?) time: (\d+) ns", p.stderr))
        t_idx = int(prof.get("indexing", 0)) * 1e-9
        want = [l for l in open(samf).read().splitlines() if not l.startswith("@")]
        file_runs["reference"] = {"wall_s": round(wall, 3), "indexing_s": round(t_idx, 3), "reads_per_s": n_reads / max(wall - t_idx, 1e-9),
                            "threads": cores, "profile_thread_seconds": {k: round(int(v) * 1e-9, 3) for k, v in prof.items()}}
        # ---- the batched C host (INTEGRATION.md level 2): the same program with genome-on-diet_b200/host/gd_batched_host.c in place
        # of map.c's pipeline -- FASTQ parsing, mapping on the GPU, SAM file written, one process, same flags
        if os.path.exists(batched_bin):
            def run_batched(kflag, env):
                samb = os.path.join(tmp, "batched.sam")
                t0 = time.perf_counter()
                pb = subprocess.run([batched_bin, "-t", str(cores), "-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.05,150,200"]
                                    + kflag + ["-o", samb, fa, fq], capture_output=True, text=True, env=dict(os.environ, GDIET_GPUS="1", **env))
                wall_b = time.perf_counter() - t0
                mm = re.search(r"\[M::mm_map_file_frag\] (\d+) reads, \d+ bases in ([0-9.]+) s", pb.stderr)
                ix = re.search(r"\[PROFILING\] indexing time: (\d+) ns", pb.stderr)
                gotb = [l for l in open(samb).read().splitlines() if not l.startswith("@")] if pb.returncode == 0 else []
                r = {"flags": kflag, "returncode": pb.returncode,
                     "wall_s": round(wall_b, 3), "indexing_s": round(int(ix.group(1)) * 1e-9, 3) if ix else None,
                     "map_pipeline_s": float(mm.group(2)) if mm else None,
                     "reads_per_s_pipeline": (n_reads / float(mm.group(2))) if mm else None,
                     "summary": (re.search(r"\[M::mm_map_file_frag\] .*", pb.stderr) or [""])[0][:300],
                     "sam_identical": gotb == want}
                if "GD_MAP_PROFILE" in env:
                    r["profile"] = [l for l in pb.stderr.splitlines() if l.startswith("[gd_")][:60]
                return r

            file_runs["batched_host"] = dict(binary="oracle/_ref/GDiet_cuda_batched_sr (unmodified reference sources + gd_batched_host.c)",
                                       mini_batch="-K 30M (200 k reads per batch)",
                                       note="map_pipeline_s = FASTQ parse + GPU mapping + SAM file write under kt_pipeline (after the index and the CUDA context exist)",
                                       **run_batched(["-K", "30M"], {}))
            if os.environ.get("SR_BATCHED_PROBE"):  # mini-batch size and the phases of one call
                file_runs["batched_host_probe"] = [run_batched(k, e) for k, e in (([], {}), (["-K", "30M"], {"GD_MAP_PROFILE": "1"}), (["-K", "150M"], {}),
                                                                             (["-K", "500M"], {}), (["-K", "500M"], {"GD_MAP_PROFILE": "1"}))]
    opt, post = gd.sr_options(), gd.sr_post_options()
    t0 = time.perf_counter()
    idx = ctx.index_build(contigs, 11, 21, "10")
    t_index = time.perf_counter() - t0
    res = {}
    for it in range(4):  # first pass warms the context's staging buffers
        l0 = ctx.stat("kernel_launches")
        t0 = time.perf_counter()
        coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, opt, cand_cap=n_reads + 1024, cigar_cap=8 * n_reads + 1024)
        t1 = time.perf_counter()
        h = gd.sr_sam_batch(names, off, lens, buf, qual, coff, cand, cig, ["chr1"], contigs, post, parts=True)  # what a host writes out
        t2 = time.perf_counter()
        sam = h.bytes() if it == 3 else b""
        h.free()
        res = dict(map_s=t1 - t0, sam_s=t2 - t1, launches=ctx.stat("kernel_launches") - l0)
    # ---- pipelined: two contexts on two host threads drive the device stage of alternate batches (their transfers and
    # read-backs overlap each other's kernels); the main thread turns finished batches into SAM text IN INPUT ORDER on
    # the remaining cores
    import threading, queue
    nb = max(1, min(8, n_reads // 125_000))
    bounds = [n_reads * b // nb for b in range(nb + 1)]
    ctx_p = gd.Context(ctx.device) if nb > 1 else None

    def producer(c, parts, q):
        for b in parts:
            lo, hi = bounds[b], bounds[b + 1]
            r = c.sr_map_batch(idx, off[lo:hi] - off[lo], lens[lo:hi], buf[off[lo]:off[lo] + (hi - lo) * 150], opt, cand_cap=(hi - lo) + 1024,
                               cigar_cap=8 * (hi - lo) + 1024)
            q.put((b, r))

    pipe_s = None
    post_p = gd.sr_post_options(n_threads=max(1, cores - 2))  # leave cores to the threads that drive the GPU
    for it in range(2):
        q = queue.Queue()
        t0 = time.perf_counter()
        ctxs = [ctx] + ([ctx_p] if ctx_p else [])
        ths = [threading.Thread(target=producer, args=(c, range(k, nb, len(ctxs)), q)) for k, c in enumerate(ctxs)]
        for t in ths:
            t.start()
        done, nxt, total_bytes = {}, 0, 0
        while nxt < nb:
            b, r = q.get()
            done[b] = r
            while nxt in done:
                co, ca, cg = done.pop(nxt)
                lo, hi = bounds[nxt], bounds[nxt + 1]
                h = gd.sr_sam_batch(C_names[nxt], off[lo:hi] - off[lo], lens[lo:hi], buf[off[lo]:off[lo] + (hi - lo) * 150],
                                    qual[off[lo]:off[lo] + (hi - lo) * 150], co, ca, cg, ["chr1"], contigs, post_p, parts=True)
                total_bytes += h.n
                h.free()
                nxt += 1
        for t in ths:
            t.join()
        pipe_s = time.perf_counter() - t0
    if ctx_p:
        ctx_p.close()
    # ---- device stage alone with TWO contexts on two host threads (one stream each, the index is shared read-only): the
    # uploads / downloads / count read-backs of one context overlap the kernels of the other
    dual_s = None
    if n_reads >= 400_000:
        ctx_b = gd.Context(ctx.device)
        nb2 = 8
        bounds2 = [n_reads * b // nb2 for b in range(nb2 + 1)]

        def worker(c, parts):
            for b in parts:
                lo, hi = bounds2[b], bounds2[b + 1]
                c.sr_map_batch(idx, off[lo:hi] - off[lo], lens[lo:hi], buf[off[lo]:off[lo] + (hi - lo) * 150], opt, cand_cap=(hi - lo) + 1024,
                               cigar_cap=8 * (hi - lo) + 1024)

        for it in range(3):
            t0 = time.perf_counter()
            ths = [threading.Thread(target=worker, args=(c, range(k, nb2, 2))) for k, c in enumerate((ctx, ctx_b))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            dual_s = time.perf_counter() - t0
        ctx_b.close()
    out = {"what": "config 1: sr end to end", "ref_bp": len(genome), "reads": n_reads, "index_build_s": round(t_index, 4),
           "index_minimizers": idx.stat("n_minimizers"), "map_batch_s": round(res["map_s"], 4), "sam_s": round(res["sam_s"], 4),
           "reads_per_s_serial": n_reads / (res["map_s"] + res["sam_s"]), "pipelined_s": round(pipe_s, 4), "pipeline_batches": nb,
           "reads_per_s": n_reads / min(pipe_s, res["map_s"] + res["sam_s"]), "reads_per_s_map_only": n_reads / res["map_s"],
           "reads_per_s_map_only_two_contexts": (n_reads / dual_s) if dual_s else None,
           "candidates": int(coff[-1]), "exact": int(cand["exact"].sum()), "gpu_launches_per_batch": res["launches"], "host_cores": cores}
    # ---- reads in, SAM text out with the post-DP stage on the device (gd_sr_map_sam_batch)
    for it in range(3):
        t0 = time.perf_counter()
        pieces = ctx.sr_map_sam_batch(idx, names, off, lens, buf, qual, opt, post, ["chr1"], join=False)
        dev_s = time.perf_counter() - t0
    import ctypes
    out["map_sam_device_s"] = round(dev_s, 4)
    out["reads_per_s_device_sam"] = n_reads / dev_s
    out["device_sam_identical_to_host_stage"] = b"".join(ctypes.string_at(a, l) for a, l in pieces) == sam
    if want is not None:
        got = sam.decode().splitlines()
        out.update(file_runs)
        out["sam_identical"] = got == want
        out["sam_lines"] = len(want)
    if want is not None:
        if got != want:
            bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
            out["sam_first_diff"] = [got[bad[0]][:300], want[bad[0]][:300]] if bad else ["length", "%d vs %d" % (len(got), len(want))]
    idx.close()
    return out


def main():
    ref_mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 5
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    run_ref = (sys.argv[3] != "noref") if len(sys.argv) > 3 else True
    ctx = gd.Context(0)
    print(json.dumps(run(ctx, ref_mbp, n_reads, run_ref)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
