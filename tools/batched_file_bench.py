#!/usr/bin/env python
"""FASTQ file in, SAM file out, whole programs side by side at a size where start-up costs no longer dominate: the unmodified
reference (`oracle/_ref/GDiet_avx_sr -t <host cores>`) against the batched C host (`oracle/_ref/GDiet_cuda_batched_sr`:
unmodified reference sources + genome-on-diet_b200/host/gd_batched_host.c) with BASELINE config 1's flags.  Reports, for
both: wall time of the process, its indexing time ([PROFILING] line of main.c), reads/s over the mapping part, and for the
batched host the pipeline line (FASTQ parse + GPU mapping + SAM write under kt_pipeline).  The two SAM files are compared line
by line (without @PG).  This process holds no device memory while the binaries run.
    python tools/batched_file_bench.py [ref_mbp=50] [n_reads=5000000] [gpus=1]"""
import hashlib, json, os, re, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
FLAGS = ["-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.05,150,200"]


def write_inputs(tmp, ref_mbp, n_reads):
    import torch
    import map_strong_bench as msb
    dev = torch.device("cuda:0")
    genome, goff, glens = msb.make_genome(dev, int(ref_mbp * 1e6), ncontig=4, seed=16)
    reads = msb.make_reads(dev, genome, goff, glens, n_reads, seed=18).cpu().numpy()
    g = genome.cpu().numpy()
    del genome
    torch.cuda.empty_cache()
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    with open(fa, "wb") as f:
        for i, (o, l) in enumerate(zip(goff, glens)):
            f.write(b">chr%d\n" % (i + 1))
            f.write(g[o:o + l].tobytes())
            f.write(b"\n")
    L = reads.shape[1]
    rec = np.empty((n_reads, 12 + L + 3 + L + 1), np.uint8)   # "@r%09d\n" + bases + "\n+\n" + qualities + "\n"
    rec[:, 0], rec[:, 1], rec[:, 11] = ord("@"), ord("r"), ord("\n")
    idx = np.arange(n_reads, dtype=np.int64)
    for d in range(9):
        rec[:, 10 - d] = ord("0") + (idx // 10 ** d) % 10
    rec[:, 12:12 + L] = reads
    rec[:, 12 + L:12 + L + 3] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 12 + L + 3:12 + 2 * L + 3] = ord("I")
    rec[:, -1] = ord("\n")
    rec.tofile(fq)
    return fa, fq


def sam_digest(path):
    h, n = hashlib.sha256(), 0
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b"@PG"):
                continue
            h.update(line)
            n += not line.startswith(b"@")
    return h.hexdigest(), n


def run_prog(binary, extra, env, fa, fq, out, cores, n_reads):
    t0 = time.perf_counter()
    p = subprocess.run([binary, "-t", str(cores)] + FLAGS + extra + ["-o", out, fa, fq], capture_output=True, text=True, env=dict(os.environ, **env))
    wall = time.perf_counter() - t0
    ix = re.search(r"\[PROFILING\] indexing time: (\d+) ns", p.stderr)
    t_idx = int(ix.group(1)) * 1e-9 if ix else 0.0
    mm = re.search(r"\[M::mm_map_file_frag\] .*", p.stderr)
    pipe = re.search(r"in ([0-9.]+) s:", mm.group(0)) if mm else None
    r = {"flags": extra, "returncode": p.returncode, "wall_s": round(wall, 3), "indexing_s": round(t_idx, 3),
         "reads_per_s_after_indexing": n_reads / max(wall - t_idx, 1e-9), "reads_per_s_wall": n_reads / wall}
    if mm:
        r["pipeline"] = mm.group(0)[:300]
        r["reads_per_s_pipeline"] = n_reads / float(pipe.group(1)) if pipe else None
    if p.returncode != 0:
        r["stderr_tail"] = p.stderr[-600:]
    if "GD_MAP_PROFILE" in env:  # the calls of the mapping step: prologue (buffers) and slices, per mini-batch
        r["profile"] = [l for l in p.stderr.splitlines() if l.startswith("[gd_sr_map_sam_batch]")][:40]
    return r


def main():
    ref_mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 50
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
    gpus = sys.argv[3] if len(sys.argv) > 3 else "1"
    cores = len(os.sched_getaffinity(0))
    tmp = tempfile.mkdtemp(prefix="gdfile_")
    fa, fq = write_inputs(tmp, ref_mbp, n_reads)
    out = {"what": "config 1 flags, file to file", "ref_bp": int(ref_mbp * 1e6), "reads": n_reads, "read_len": 150, "host_cores": cores,
           "fastq_bytes": os.path.getsize(fq), "gpus": int(gpus)}
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "GDiet_avx_sr")
    bat_bin = os.path.join(ROOT, "oracle", "_ref", "GDiet_cuda_batched_sr")
    ref_sam, bat_sam = os.path.join(tmp, "ref.sam"), os.path.join(tmp, "batched.sam")
    out["reference"] = run_prog(ref_bin, [], {}, fa, fq, ref_sam, cores, n_reads)
    want, n_want = sam_digest(ref_sam)
    out["sam_records"] = n_want
    out["batched_host"] = []
    for extra in ([], ["-K", "150M"], ["-K", "500M"]):
        r = run_prog(bat_bin, extra, {"GDIET_GPUS": gpus}, fa, fq, bat_sam, cores, n_reads)
        got, n_got = sam_digest(bat_sam) if r["returncode"] == 0 else ("", 0)
        r["sam_file_identical"] = got == want and n_got == n_want
        out["batched_host"].append(r)
    if os.environ.get("FILE_BENCH_PROFILE"):
        out["batched_host_profile"] = run_prog(bat_bin, [], {"GDIET_GPUS": gpus, "GD_MAP_PROFILE": "1"}, fa, fq, bat_sam, cores, n_reads)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
