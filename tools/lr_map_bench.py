#!/usr/bin/env python
"""BASELINE configs 3 / 4 (long reads), device mapping stage: `-ax map-hifi -Z 10 -W 2 -k 19 -w 19 -r 1000` on 15 kbp
HiFi-like reads (1 % error) or `-ax map-ont -Z 10 -W 2 -k 15 -w 10 -r 1300` + the README voting flags on 50 kbp ONT-like
reads (8 % error), against a synthetic reference.  gd_lr_map_batch does everything of LR/map.c:mm_map_frag up to the
ksw_extz_t of every candidate on the GPU; the CIGAR stitching (concatenate_cigars) and SAM stay with the host program.
Parity: the first `n_check` reads are compared with the call trace of the unmodified reference program (-t 1); the CPU
baseline is the same program on all host cores over all reads ([PROFILING] thread-seconds and wall time).
    python tools/lr_map_bench.py hifi|ont|ont-asis [ref_mbp] [n_reads] [n_check]"""
import json, os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gdiet_b200 as gd
from gdiet_b200 import synth

ONT_FLAGS = ["--vt_dis=1000", "--vt_nb_loc=3", "--vt_df1=0.007", "--vt_df2=0.007", "--max_min_gap=4000", "--vt_f=0.04", "--vt_cov", "0.3",
             "--sort=merge", "--frag=no"]


def band_cells(ql, tl, w):
    """banded cells of one ksw_extd2 call: sum over anti-diagonals of en0 - st0 + 1"""
    r = np.arange(ql + tl - 1, dtype=np.int64)
    st = np.maximum(np.maximum(0, r - ql + 1), (r - w + 1) >> 1)
    en = np.minimum(np.minimum(tl - 1, r), (r + w) >> 1)
    return int(np.maximum(en - st + 1, 0).sum())


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "hifi"
    ref_mbp = float(sys.argv[2]) if len(sys.argv) > 2 else 100
    n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else (2000 if kind == "hifi" else 500)
    n_check = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    ctx = gd.Context(0)
    print(json.dumps(run(ctx, kind, ref_mbp, n_reads, n_check)), flush=True)
    ctx.close()


def run(ctx, kind="hifi", ref_mbp=100, n_reads=2000, n_check=100, run_ref=True):
    if kind == "hifi":
        preset, k, w, bw, L, sub, indel, extra, okw = "map-hifi", 19, 19, 1000, 15000, 0.005, 0.005, [], {}
    elif kind == "ont":  # BASELINE config 4's flags (-r 1300 -s 35000) + the README's voting flags (SURVEY.md finding 6)
        preset, k, w, bw, L, sub, indel, extra = "map-ont", 15, 10, 1300, 50000, 0.03, 0.05, ONT_FLAGS + ["-s", "35000"]
        okw = dict(vt_dis=1000, vt_df1=0.007, vt_df2=0.007, vt_f=0.04, vt_cov=0.3)
    else:  # "ont-asis": config 4 exactly as BASELINE writes it -- the long-read tree's default voting thresholds reject every read
        preset, k, w, bw, L, sub, indel, extra, okw = "map-ont", 15, 10, 1300, 50000, 0.03, 0.05, ["-s", "35000"], {}
    cores = len(os.sched_getaffinity(0))
    rng = np.random.default_rng(5)
    ncontig = 4
    contigs = [synth.random_genome(int(ref_mbp * 1e6) // ncontig, seed=40 + i) for i in range(ncontig)]
    lut = np.zeros(256, np.uint8)
    lut[synth.ACGTN] = np.arange(5)
    reads = []
    for i in range(n_reads):
        c = contigs[int(rng.integers(0, ncontig))]
        st = int(rng.integers(0, len(c) - int(L * 1.2)))
        codes = synth.mutate_codes(rng, lut[c[st:st + int(L * 1.2)]], sub + indel, sub=sub / (sub + indel), dele=indel / (sub + indel) / 2)[:L]
        if rng.random() < 0.5:
            codes = (3 - codes)[::-1]
        reads.append(synth.ACGTN[codes])
    lens = np.array([len(r) for r in reads], np.int32)
    off = np.zeros(n_reads, np.int64)
    off[1:] = np.cumsum(lens[:-1].astype(np.int64))
    buf = np.concatenate(reads)
    # ---- whole programs, file to file, BEFORE this process takes device memory for its own contexts (the DP backtrack arena is sized
    # from what is free: a second process next to a loaded one would get small launches): the unmodified reference on all host
    # cores, then the batched C host of the long-read tree (INTEGRATION.md level 2) with the same flags
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "GDiet_avx_lr")
    flags = ["-ax", preset, "-Z", "10", "-W", "2", "-k", str(k), "-w", str(w), "-r", str(bw)] + extra
    file_runs = {}
    if run_ref and os.path.exists(ref_bin):
        import maplib
        tmp = tempfile.mkdtemp(prefix="gdref_")
        fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
        maplib.write_fasta(fa, contigs)
        maplib.write_fastq(fq, reads)
        t0 = time.perf_counter()
        p = subprocess.run([ref_bin, "-t", str(cores)] + flags + ["-o", os.path.join(tmp, "out.sam"), fa, fq], capture_output=True, text=True)
        wall = time.perf_counter() - t0
        prof = dict(re.findall(r"\[PROFILING\] (.+?) time: (\d+) ns", p.stderr))
        t_idx = int(prof.get("indexing", 0)) * 1e-9
        # ---- the batched C host of the long-read tree (INTEGRATION.md level 2): FASTQ file in, SAM file out, same flags
        batched_bin = os.path.join(ROOT, "oracle", "_ref", "GDiet_cuda_batched_lr")
        if os.path.exists(batched_bin):
            strip = lambda path: [l for l in open(path).read().splitlines() if not l.startswith("@PG")]
            want_sam = strip(os.path.join(tmp, "out.sam"))

            def run_batched(extra, env, prefix=()):
                samb = os.path.join(tmp, "batched.sam")
                t0 = time.perf_counter()
                pb = subprocess.run(list(prefix) + [batched_bin, "-t", str(cores)] + flags + extra + ["-o", samb, fa, fq], capture_output=True, text=True,
                                    env=dict(os.environ, GDIET_GPUS="1", **env))
                wall_b = time.perf_counter() - t0
                mm = re.search(r"\[M::mm_map_file_frag\] (\d+) reads, \d+ bases in ([0-9.]+) s.*", pb.stderr)
                ix = re.search(r"\[PROFILING\] indexing time: (\d+) ns", pb.stderr)
                r = {"flags": extra, "env": env, "returncode": pb.returncode, "wall_s": round(wall_b, 2),
                     "indexing_s": round(int(ix.group(1)) * 1e-9, 2) if ix else None, "map_pipeline_s": float(mm.group(2)) if mm else None,
                     "reads_per_s_pipeline": (n_reads / float(mm.group(2))) if mm else None,
                     "summary": mm.group(0)[:300] if mm else pb.stderr[-300:],
                     "sam_file_identical": pb.returncode == 0 and strip(samb) == want_sam}
                r["log"] = [l[:160] for l in pb.stderr.splitlines() if l.startswith("[M::") and "mapped" not in l][:24]  # (time stamps of the phases)
                if "GD_MAP_PROFILE" in env:
                    r["profile"] = [l for l in pb.stderr.splitlines() if l.startswith("[gd_")][:80]
                return r

            file_runs["batched_host"] = dict(binary="oracle/_ref/GDiet_cuda_batched_lr (unmodified reference sources + gd_batched_host.c)",
                                       **run_batched([], {}))
            if os.environ.get("LR_BATCHED_NCU"):  # per-kernel durations of the same command (launch list; not a timing run)
                file_runs["batched_host_ncu"] = run_batched([], {}, ("ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv",
                                                               "--log-file", os.environ["LR_BATCHED_NCU"]))
            if os.environ.get("LR_BATCHED_PROBE"):  # where the time of the file-to-file pipeline goes: mini-batch size, device slices
                eager = {"CUDA_MODULE_LOADING": "EAGER"}
                file_runs["batched_host_probe"] = [run_batched(f, e) for f, e in (
                    ([], {}), ([], eager), ([], {}), ([], eager), ([], {"GD_MAP_PROFILE": "1"}), ([], dict(eager, GD_MAP_PROFILE="1")))]
    ctx.set_option("time_kernels", 1)
    t0 = time.perf_counter()
    idx = ctx.index_build(contigs, w, k, "10")
    t_index = time.perf_counter() - t0
    lo, hi = (50, 500) if preset == "map-hifi" else (10, 1000000)
    mid = min(max(idx.cal_max_occ(2e-4), lo), hi)
    o = gd.lr_options(preset, bw=bw, mid_occ=mid, **okw)
    tm = []
    for it in range(3):
        ctx.stat("ksw_dp_reset")
        t0 = time.perf_counter()
        coff, cand, cig = ctx.lr_map_batch(idx, off, lens, buf, o, cand_cap=6 * n_reads, cigar_cap=max(1 << 22, int(lens.sum()) // 4))
        tm.append(time.perf_counter() - t0)
        dp_us = ctx.stat("ksw_dp_us")
    cells = sum(band_cells(int(c["qe"] - c["qs"]), int(c["re"] - c["rs"]), bw) for c in cand if not c["exact"])
    out = {"what": "config %s: long-read mapping stage on the device" % ({"hifi": "3 (hifi)", "ont": "4 (ont, -s 35000 + README voting flags)"}.get(kind, "4 exactly as written (ont, -s 35000, default voting thresholds: nothing maps)")),
           "ref_bp": int(sum(len(c) for c in contigs)), "reads": n_reads, "read_len": L, "band": bw, "mid_occ": mid, "index_build_s": round(t_index, 3),
           "map_batch_s": [round(x, 3) for x in tm], "reads_per_s": n_reads / min(tm), "bases_per_s": float(lens.sum()) / min(tm),
           "candidates": int(coff[-1]), "mapped_reads": int((np.diff(coff) > 0).sum()), "chained": int((cand["reserved"][:, 0] >= 0).sum()),
           "dp_cells": cells, "dp_kernel_ms": dp_us / 1e3, "dp_kernel_gcups": cells / (dp_us * 1e-6) / 1e9 if dp_us else None,
           "stage_gcups": cells / min(tm) / 1e9, "host_cores": cores}
    # ---- host stage: post-processing + SAM records (reads whose candidates need CIGAR stitching are flagged, not written)
    names = gd._cstr_array(["r%d" % i for i in range(n_reads)])
    qual = np.full(len(buf), ord("I"), np.uint8)
    post = gd.lr_post_options(preset)
    if kind != "hifi":
        post.min_dp_max = 35000  # -s 35000
    ref = gd.flat_ref(contigs)  # (a C host holds the reference as one buffer already)
    for it in range(2):
        t0 = time.perf_counter()
        sam_txt, sam_off, stitch = gd.lr_sam_batch(names, off, lens, buf, qual, coff, cand, cig, ["chr%d" % (i + 1) for i in range(ncontig)], contigs, post,
                                                   ref=ref)
        out["sam_s"] = round(time.perf_counter() - t0, 3)
    out["reads_needing_stitch"] = int(stitch.sum())
    out["reads_per_s_end_to_end"] = n_reads / (min(tm) + out["sam_s"])
    if run_ref and os.path.exists(ref_bin):
        if n_check > 0:
            _, tr = maplib.run_reference(contigs, reads[:n_check], flags, program=ref_bin, threads=1)
            bad = 0
            for i, t in enumerate(tr):
                try:
                    maplib.lr_cands_equal_trace(cand[coff[i]:coff[i + 1]], cig, t["cands"], "read %d" % i)
                except AssertionError as e:
                    bad += 1
                    if bad <= 3:
                        print(str(e)[:300], file=sys.stderr)
            out["parity"] = {"reads_checked": len(tr), "dp_calls_checked": int(sum(len(t["cands"]) for t in tr)), "mismatching_reads": bad}
        want = {}
        for l in open(os.path.join(tmp, "out.sam")).read().splitlines():
            if not l.startswith("@"):
                want.setdefault(l.split("\t", 1)[0], []).append(l)
        same = diff = 0
        for i in range(n_reads):
            if stitch[i]:
                continue
            mine = sam_txt[sam_off[i]:sam_off[i + 1]].decode().splitlines()
            if mine == want.get("r%d" % i, []):
                same += 1
            else:
                diff += 1
        out["sam"] = {"reads_identical": same, "reads_different": diff, "reads_left_to_host_stitching": int(stitch.sum())}
        out.update(file_runs)
        out["reference"] = {"wall_s": round(wall, 2), "indexing_s": round(t_idx, 2), "reads_per_s": n_reads / max(wall - t_idx, 1e-9), "threads": cores,
                            "profile_thread_seconds": {kk: round(int(v) * 1e-9, 2) for kk, v in prof.items()}}
    idx.close()
    ctx.set_option("time_kernels", 0)
    return out


if __name__ == "__main__":
    main()
