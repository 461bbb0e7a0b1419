"""Level-0 drop-in: the UNMODIFIED reference program, compiled from /root/reference by oracle/Makefile
without sketch.c and ksw2_extd2_avx.c and linked against libgdiet_cuda.so instead (oracle/_ref/GDiet_cuda_sr|lr), so that
its mm_sketch / mm_sketch2 / mm_sketch3 / ksw_extd2_avx512 calls run on the GPU through the library's drop-in symbols.
Its SAM output must equal the SAM of the all-CPU build (oracle/_ref/GDiet_avx_sr|lr) byte for byte."""
import os, re
import subprocess
import tempfile

import pytest

import maplib
from oraclelib import ORACLE_DIR, cpu_has_avx512

pytestmark = pytest.mark.gpu

CUDA_SR = os.path.join(ORACLE_DIR, "_ref", "GDiet_cuda_sr")
CUDA_LR = os.path.join(ORACLE_DIR, "_ref", "GDiet_cuda_lr")


def run(prog, flags, fa, fq, out, threads):
    p = subprocess.run([prog, "-t", str(threads)] + flags + ["-o", out, fa, fq], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in open(out).read().splitlines() if not l.startswith("@PG")]


@pytest.mark.skipif(not (os.path.exists(CUDA_SR) and maplib.have_ref_program() and cpu_has_avx512()),
                    reason="needs oracle/_ref/GDiet_cuda_sr + GDiet_avx_sr (built where /root/reference exists)")
def test_reference_program_on_the_drop_in_symbols_short_reads():
    contigs, reads = maplib.make_dataset(seed=51, n_reads=1500)
    tmp = tempfile.mkdtemp(prefix="gddrop_")
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    flags = ["-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.05,150,200", "-n", "0.2,0.1"]
    want = run(maplib.REF_SR, flags, fa, fq, os.path.join(tmp, "cpu.sam"), 2)
    got = run(CUDA_SR, flags, fa, fq, os.path.join(tmp, "gpu.sam"), 2)
    assert len(got) == len(want) and got == want


@pytest.mark.skipif(not (os.path.exists(CUDA_LR) and os.path.exists(maplib.REF_LR) and cpu_has_avx512()),
                    reason="needs oracle/_ref/GDiet_cuda_lr + GDiet_avx_lr")
def test_reference_program_on_the_drop_in_symbols_long_reads():
    contigs, reads = maplib.make_long_dataset(seed=52, read_len=8000, n_reads=40)
    tmp = tempfile.mkdtemp(prefix="gddrop_")
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    flags = ["-ax", "map-hifi", "-Z", "10", "-W", "2", "-k", "19", "-w", "19", "-r", "600"]
    want = run(maplib.REF_LR, flags, fa, fq, os.path.join(tmp, "cpu.sam"), 2)
    got = run(CUDA_LR, flags, fa, fq, os.path.join(tmp, "gpu.sam"), 2)
    assert len(got) == len(want) and got == want


# ---- Level 2: the batched host (genome-on-diet_b200/host/gd_batched_host.c) in place of map.c's pipeline ----------------
BATCHED_SR = os.path.join(ORACLE_DIR, "_ref", "GDiet_cuda_batched_sr")
BATCHED_LR = os.path.join(ORACLE_DIR, "_ref", "GDiet_cuda_batched_lr")


def n_gpus():
    import torch
    return torch.cuda.device_count()


def run_env(prog, flags, fa, fq, out, threads, env):
    p = subprocess.run([prog, "-t", str(threads)] + flags + ["-o", out, fa, fq], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **env))
    assert p.returncode == 0, p.stderr[-2000:]
    # the device pipeline ran, and the reference's CPU mapping code (still linked, under another name) did not: its [PROFILING]
    # counters stay at zero
    assert "[M::mm_map_file_frag]" in p.stderr and "step seconds" in p.stderr, p.stderr[-2000:]
    for m in re.finditer(r"\[PROFILING\] (seeding|voting|sequence alignment|pattern alignment) time: (\d+) ns", p.stderr):
        assert int(m.group(2)) == 0, m.group(0)
    return [l for l in open(out).read().splitlines() if not l.startswith("@PG")], p.stderr


@pytest.mark.skipif(not (os.path.exists(BATCHED_SR) and maplib.have_ref_program() and cpu_has_avx512()),
                    reason="needs oracle/_ref/GDiet_cuda_batched_sr + GDiet_avx_sr (built where /root/reference exists)")
@pytest.mark.parametrize("flags,read_len,ragged", [
    (["-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.05,150,200"], 150, False),             # BASELINE config 1
    (["-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.25,20,60", "-n", "0.2,0.1"], 300, True),  # band per read
])
def test_batched_host_short_reads_sam_identical(flags, read_len, ragged):
    """FASTQ in, SAM file out through the batched C host (reader -> gd_multi_sr_map_sam -> writer under kt_pipeline), with a
    mini-batch size (-K) that cuts the input into several batches: byte-identical to GDiet_avx"""
    import numpy as np
    contigs, reads = maplib.make_dataset(seed=61, n_reads=6000, read_len=read_len)
    if ragged:
        rng = np.random.default_rng(3)
        reads = [r[:int(rng.integers(60, read_len + 1))].copy() for r in reads]
    tmp = tempfile.mkdtemp(prefix="gdbatch_")
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    want = run(maplib.REF_SR, flags, fa, fq, os.path.join(tmp, "cpu.sam"), 2)
    got, _ = run_env(BATCHED_SR, flags + ["-K", "200k"], fa, fq, os.path.join(tmp, "gpu.sam"), 3, {"GDIET_GPUS": "1"})
    assert len(got) == len(want) and got == want
    if n_gpus() >= 2:  # read shards on two devices, index broadcast, SAM pieces in input order
        for env in ({"GDIET_GPUS": "2"}, {"GDIET_GPUS": "2", "GDIET_NO_NCCL": "1"}):
            got2, err = run_env(BATCHED_SR, flags + ["-K", "200k"], fa, fq, os.path.join(tmp, "gpu2.sam"), 4, env)
            assert got2 == want, err[-1000:]


@pytest.mark.skipif(not (os.path.exists(BATCHED_SR) and maplib.have_ref_program() and cpu_has_avx512()),
                    reason="needs oracle/_ref/GDiet_cuda_batched_sr + GDiet_avx_sr (built where /root/reference exists)")
def test_batched_host_index_paths():
    """Where the batched host's index comes from: (a) a FASTA reference -- mapped and parsed by the host file (a gzip one: read
    through the reference's reader), index built on the device, no host hash tables (the statistics line says so); (b) the same forced through the reference's mm_idx_gen (GDIET_REF_INDEX);
    (c) an .mmi written by the reference program (-d): loaded by the reference's code, device index built from the sequences
    it holds; (d) a reference cut into several index parts (-I): every part is mapped in turn as main.c loops.  SAM identical
    to GDiet_avx each time."""
    contigs, reads = maplib.make_dataset(seed=64, n_reads=3000, read_len=150)  # three contigs, 300 / 200 / 100 kbp
    tmp = tempfile.mkdtemp(prefix="gdbatch_")
    fa, fq, mmi = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq"), os.path.join(tmp, "ref.mmi")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    flags = ["-ax", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-r", "0.05,150,200"]
    want = run(maplib.REF_SR, flags, fa, fq, os.path.join(tmp, "cpu.sam"), 2)
    got, err = run_env(BATCHED_SR, flags, fa, fq, os.path.join(tmp, "a.sam"), 3, {"GDIET_GPUS": "1"})
    assert got == want and "(device index) distinct minimizers" in err and "device index on 1 GPU" in err, err[-1500:]
    got, err = run_env(BATCHED_SR, flags, fa, fq, os.path.join(tmp, "b.sam"), 3, {"GDIET_GPUS": "1", "GDIET_REF_INDEX": "1"})
    assert got == want and "(device index) distinct minimizers" not in err and "device index on 1 GPU" in err, err[-1500:]
    # the same reference with 60-column lines, CRLF in one record, blank lines and an empty record, parsed record-wise on several threads
    fa60 = os.path.join(tmp, "ref60.fa")
    with open(fa60, "wb") as f:
        for i, c in enumerate(contigs):
            nl = b"\r\n" if i == 1 else b"\n"
            f.write(b">chr%d some description" % (i + 1) + nl)
            b = bytes(c)
            f.write(b"".join(b[j:j + 60] + nl for j in range(0, len(b), 60)))
            if i == 0:
                f.write(b"\n\n")
    got, err = run_env(BATCHED_SR, flags, fa60, fq, os.path.join(tmp, "p.sam"), 4, {"GDIET_GPUS": "1", "GDIET_REF_PAR_MIN_BYTES": "0"})
    assert got == want and "records parsed on several threads" in err, err[-1500:]
    got, err = run_env(BATCHED_SR, flags, fa60, fq, os.path.join(tmp, "p1.sam"), 4, {"GDIET_GPUS": "1", "GDIET_REF_ONE_THREAD": "1"})
    assert got == want and "records parsed on several threads" not in err, err[-1500:]
    import gzip
    with open(fa, "rb") as f, gzip.open(fa + ".gz", "wb", compresslevel=1) as g:  # a gzip reference goes through the reference's reader
        g.write(f.read())
    got, err = run_env(BATCHED_SR, flags, fa + ".gz", fq, os.path.join(tmp, "gz.sam"), 3, {"GDIET_GPUS": "1"})
    assert got == want and "(device index) distinct minimizers" in err, err[-1500:]
    p = subprocess.run([maplib.REF_SR, "-t", "2", "-x", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", "-d", mmi, fa],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and os.path.getsize(mmi) > 1000, p.stderr[-1000:]
    want_mmi = run(maplib.REF_SR, flags, mmi, fq, os.path.join(tmp, "cpu_mmi.sam"), 2)
    got, err = run_env(BATCHED_SR, flags, mmi, fq, os.path.join(tmp, "c.sam"), 3, {"GDIET_GPUS": "1"})
    assert got == want_mmi and "device index on 1 GPU" in err, err[-1500:]
    part = ["-I", str(max(len(c) for c in contigs) + 10)]  # every contig its own part (a part closes once it EXCEEDS -I bases)
    want_parts = run(maplib.REF_SR, flags + part, fa, fq, os.path.join(tmp, "cpu_parts.sam"), 2)
    got, err = run_env(BATCHED_SR, flags + part, fa, fq, os.path.join(tmp, "d.sam"), 3, {"GDIET_GPUS": "1"})
    assert got == want_parts and err.count("device index on 1 GPU") >= 2, err[-1500:]


@pytest.mark.skipif(not (os.path.exists(BATCHED_LR) and os.path.exists(maplib.REF_LR) and cpu_has_avx512()),
                    reason="needs oracle/_ref/GDiet_cuda_batched_lr + GDiet_avx_lr")
def test_batched_host_long_reads_sam_identical():
    """BASELINE config 3 shape (HiFi-like reads, -ax map-hifi -r 1000) through the batched C host of the long-read tree"""
    contigs, reads = maplib.make_long_dataset(seed=62, read_len=9000, n_reads=60)
    tmp = tempfile.mkdtemp(prefix="gdbatch_")
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    flags = ["-ax", "map-hifi", "-Z", "10", "-W", "2", "-k", "19", "-w", "19", "-r", "1000"]
    want = run(maplib.REF_LR, flags, fa, fq, os.path.join(tmp, "cpu.sam"), 2)
    got, _ = run_env(BATCHED_LR, flags + ["-K", "200k"], fa, fq, os.path.join(tmp, "gpu.sam"), 3, {"GDIET_GPUS": "1"})
    assert len(got) == len(want) and got == want
    # the long-read host also closes a mini-batch once it holds enough work for the devices (here: 7 reads and 50 kbases)
    got3, err3 = run_env(BATCHED_LR, flags, fa, fq, os.path.join(tmp, "gpu3.sam"), 3,
                         {"GDIET_GPUS": "1", "GDIET_LR_BATCH_BASES": "50000", "GDIET_LR_BATCH_READS": "7"})
    assert got3 == want and len(re.findall(r"mapped \d+ sequences", err3)) >= 6, err3[-1000:]
    if n_gpus() >= 2:
        got2, err = run_env(BATCHED_LR, flags + ["-K", "200k"], fa, fq, os.path.join(tmp, "gpu2.sam"), 4, {"GDIET_GPUS": "2"})
        assert got2 == want, err[-1000:]


@pytest.mark.skipif(not (os.path.exists(BATCHED_LR) and os.path.exists(maplib.REF_LR) and cpu_has_avx512()),
                    reason="needs oracle/_ref/GDiet_cuda_batched_lr + GDiet_avx_lr")
def test_batched_host_ont_flags_sam_identical():
    """BASELINE config 4's flag set (-ax map-ont -r 1300 -s ... + the README's voting flags) through the batched C host: every
    vt_* option, --max_min_gap, --sort=merge, --frag=no and -s reach the device stage and the host SAM stage as the reference's
    mm_mapopt_t holds them; reads with structural variation so that chained candidates are stitched."""
    contigs, reads = maplib.make_long_dataset(seed=63, read_len=9000, sub=0.02, indel=0.03, n_reads=80, sv_frac=0.5)
    tmp = tempfile.mkdtemp(prefix="gdbatch_")
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    flags = ["-ax", "map-ont", "-Z", "10", "-W", "2", "-k", "15", "-w", "10", "-r", "800", "-s", "3000", "--vt_dis=1000", "--vt_nb_loc=3",
             "--vt_df1=0.007", "--vt_df2=0.007", "--max_min_gap=4000", "--vt_f=0.04", "--vt_cov", "0.3", "--sort=merge", "--frag=no"]
    want = run(maplib.REF_LR, flags, fa, fq, os.path.join(tmp, "cpu.sam"), 2)
    got, _ = run_env(BATCHED_LR, flags, fa, fq, os.path.join(tmp, "gpu.sam"), 3, {"GDIET_GPUS": "1"})
    mapped = sum(1 for l in want if not l.startswith("@") and l.split("\t")[2] != "*")
    assert got == want and mapped >= 20, (len(got), len(want), mapped)


def test_batched_host_refuses_what_the_device_path_does_not_cover():
    if not os.path.exists(BATCHED_SR):
        pytest.skip("needs oracle/_ref/GDiet_cuda_batched_sr")
    contigs, reads = maplib.make_dataset(seed=63, n_reads=20)
    tmp = tempfile.mkdtemp(prefix="gdbatch_")
    fa, fq = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "reads.fq")
    maplib.write_fasta(fa, contigs)
    maplib.write_fastq(fq, reads)
    p = subprocess.run([BATCHED_SR, "-x", "sr", "-Z", "10", "-W", "2", "-k", "21", "-w", "11", fa, fq], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "PAF output is not covered" in p.stderr   # no -a: an error, not a different answer
