"""CPU: the mini-batch reader of the batched C host (genome-on-diet_b200/host/gd_batched_host.c: mapped file, memchr, bases
written straight into the pinned batch buffers) against the reference's own reader (mm_bseq_read3 over kseq.h) on files
that exercise kseq's rules: multi-line records, CRLF, empty lines, names with comments, '@' and '>' as first quality
characters, U bases, lower case, a missing final newline, a truncated last record, and several -K sizes."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oraclelib import ORACLE_DIR

CHECK = os.path.join(ORACLE_DIR, "_ref", "reader_check")
pytestmark = pytest.mark.skipif(not os.path.exists(CHECK), reason="needs oracle/_ref/reader_check (built where /root/reference exists)")


def run(text, k, threads=1):
    with tempfile.NamedTemporaryFile(suffix=".fq", delete=False) as f:
        f.write(text)
    try:
        p = subprocess.run([CHECK, f.name, str(k), str(threads)], capture_output=True, text=True, timeout=120)
        return p.returncode, p.stdout.strip() + p.stderr.strip()
    finally:
        os.unlink(f.name)


def fastq(rng, n, max_len=300, crlf=False, wrap=0, blank=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        L = int(rng.integers(1, max_len))
        seq = bytes(rng.choice(list(b"ACGTNacgtUu"), L).astype(np.uint8))
        qual = bytes(rng.integers(33, 74, L).astype(np.uint8))
        if i % 7 == 0:
            qual = b"@" + qual[1:]  # a quality line may start with '@' ...
        if i % 11 == 0:
            qual = b">" + qual[1:]  # ... or '>'
        name = b"read%d" % i + (b" comment with spaces" if i % 3 == 0 else b"\tx" if i % 5 == 0 else b"")
        def lines(s):
            if not wrap:
                return s + nl
            return b"".join(s[j:j + wrap] + nl for j in range(0, len(s), wrap))
        out.append(b"@" + name + nl + lines(seq) + b"+" + (name if i % 4 == 0 else b"") + nl + lines(qual) + (nl if blank and i % 6 == 0 else b""))
    return b"".join(out)


def fasta(rng, n, wrap=60, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        L = int(rng.integers(1, 700))
        seq = bytes(rng.choice(list(b"ACGTNacgtu"), L).astype(np.uint8))
        out.append(b">ctg%d desc" % i + nl + b"".join(seq[j:j + wrap] + nl for j in range(0, L, wrap)) + (nl if i % 4 == 0 else b""))
    return b"".join(out)


@pytest.mark.parametrize("k", [1, 500, 5000, 10 ** 9])
def test_reader_equals_reference_reader(k):
    rng = np.random.default_rng(k % 97)
    cases = [fastq(rng, 400), fastq(rng, 300, crlf=True), fastq(rng, 200, wrap=50), fastq(rng, 200, wrap=37, crlf=True),
             fastq(rng, 200, blank=True), fasta(rng, 150), fasta(rng, 100, wrap=80, crlf=True),
             fastq(rng, 50)[:-1],                       # no final newline
             b"junk before the first record\n" + fastq(rng, 20),
             b"@only_name\n", b"@a\nACGT\n+\nII\n@b\nAC\n+\nII\n",  # truncated quality: kseq's -2 ends the batch, the next call resynchronises
             b">x\n\n\nACGT\n\nAC\n>y\n>z\nA\n", b"", b"\n\n", b"@r\r\nAC\r\n+\r\nII\r\n@s\nA\r\n+\nI\r\n"]
    for i, txt in enumerate(cases):
        for threads in (1, 4):  # 4: batches of strict four-line FASTQ go through the several-thread reader, the rest falls back
            rc, out = run(txt, k, threads)
            assert rc == 0 and out.startswith("OK"), "case %d -K %d threads %d: %s" % (i, k, threads, out)


def strict_fastq(rng, n, max_len=300, min_len=1, at_qual=True):
    out = []
    for i in range(n):
        L = int(rng.integers(min_len, max_len))
        seq = bytes(rng.choice(list(b"ACGTNacgtUu"), L).astype(np.uint8))
        qual = bytes(rng.integers(33, 74, L).astype(np.uint8))
        if at_qual and i % 3 == 0:
            qual = b"@" + qual[1:]   # looks like a header to a naive splitter
        if at_qual and i % 5 == 0:
            qual = b"+" + qual[1:]
        name = b"r%d" % i + (b" some comment" if i % 4 == 0 else b"")
        out.append(b"@" + name + b"\n" + seq + b"\n+" + (name if i % 6 == 0 else b"") + b"\n" + qual + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("k", [300, 20000, 10 ** 9])
@pytest.mark.parametrize("threads", [2, 4, 16])
def test_several_thread_reader_on_strict_fastq(k, threads):
    """Strict four-line FASTQ is read by the several-thread path (the checker reports how many batches took it), with quality
    lines that start with '@' or '+' at the places where the byte range is cut; same batches, names, bases and qualities as
    the reference's reader."""
    rng = np.random.default_rng(k % 89 + threads)
    for n, txt in ((3000, strict_fastq(rng, 3000)), (500, strict_fastq(rng, 500, max_len=4000, min_len=3000)),
                   (2000, strict_fastq(rng, 2000, max_len=3)), (700, strict_fastq(rng, 700)[:-1])):
        rc, out = run(txt, k, threads)
        assert rc == 0 and out.startswith("OK"), out
        batches, reads, par = map(int, out.split()[1:4])
        assert par >= 1 and reads == n and par >= batches // 2, out


def test_several_thread_reader_falls_back_inside_a_file():
    """A file that is strict FASTQ except for one record in the middle (multi-line, CRLF, a blank line, an empty name, or an empty
    sequence): the batches around it still match the reference's reader."""
    rng = np.random.default_rng(3)
    head, tail = strict_fastq(rng, 400), strict_fastq(rng, 400, at_qual=False)  # (after a truncated record the resynchronisation may land on any "@")
    for odd in (b"@m\nACGT\nACGT\n+\nIIII\nIIII\n", b"@c\r\nACGT\r\n+\r\nIIII\r\n", b"\n",
                b"@\nAC\n+\nII\n", b"@e\n\n+\n\n"):
        for k in (2000, 10 ** 9):
            rc, out = run(head + odd + tail, k, 4)
            assert rc == 0 and out.startswith("OK"), "%r -K %d: %s" % (odd, k, out)
