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


def run(text, k):
    with tempfile.NamedTemporaryFile(suffix=".fq", delete=False) as f:
        f.write(text)
    try:
        p = subprocess.run([CHECK, f.name, str(k)], capture_output=True, text=True, timeout=120)
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
        rc, out = run(txt, k)
        assert rc == 0 and out.startswith("OK"), "case %d -K %d: %s" % (i, k, out)
