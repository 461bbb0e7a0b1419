"""Level-0 drop-in: the UNMODIFIED reference program, compiled from /root/reference by oracle/Makefile
without sketch.c and ksw2_extd2_avx.c and linked against libgdiet_cuda.so instead (oracle/_ref/GDiet_cuda_sr|lr), so that
its mm_sketch / mm_sketch2 / mm_sketch3 / ksw_extd2_avx512 calls run on the GPU through the library's drop-in symbols.
Its SAM output must equal the SAM of the all-CPU build (oracle/_ref/GDiet_avx_sr|lr) byte for byte."""
import os
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
