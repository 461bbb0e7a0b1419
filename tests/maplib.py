"""Checker-side helpers for the short-read mapping stage: synthetic datasets,
the traced reference program (oracle/_ref/GDiet_avx_sr + oracle/ref_trace.c) and the ctypes binding of
oracle/gd_oracle_map.c.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np

from oraclelib import ORACLE_DIR, Oracle

import gdiet_b200 as gd
from gdiet_b200 import synth

REF_SR = os.path.join(ORACLE_DIR, "_ref", "GDiet_avx_sr")
REF_LR = os.path.join(ORACLE_DIR, "_ref", "GDiet_avx_lr")

CAND_FIELDS = ["rid", "rs", "re", "qs", "qe", "rev", "votes", "first_q", "last_q", "exact", "score", "n_cigar", "cigar_off"]
CAND_DTYPE = np.dtype([(f, np.int32) for f in CAND_FIELDS] + [("reserved", np.int32, 3)])
DBG_DTYPE = np.dtype([(f, np.uint32) for f in
                      ["shift", "tmp_extracted_len", "n_mv", "n_a_for", "n_a_rev", "vt_threshold", "nb_potentials", "reserved"]])


SrOpt = gd.gd_sr_opt_t   # gdo_sr_opt_t of oracle/gd_oracle_map.h has the same layout
sr_opt = gd.sr_options


def ref_cmdline(o, k=21, w=11, extra=()):
    """Command-line flags of the reference program that produce the options `o` (sr preset)."""
    return ["-ax", "sr", "-Z", o.Z.decode(), "-W", str(o.W), "-k", str(k), "-w", str(w)] + list(extra)


# ------------------------------------------------------------------------------------------------
# datasets
# ------------------------------------------------------------------------------------------------
def make_dataset(seed=1, contig_lens=(300000, 200000, 100000), n_reads=3000, read_len=150, sub=0.01, indel=0.004,
                 n_frac=0.002, junk_frac=0.03, repeat=True):
    """Multi-contig reference with a planted 2 kbp three-copy repeat and reads with substitutions, real
    indels, a few N and a few unmappable random reads; reads near contig ends are included."""
    rng = np.random.default_rng(seed)
    contigs = [synth.random_genome(n, seed=seed * 100 + i) for i, n in enumerate(contig_lens)]
    if repeat and len(contigs) >= 2:
        src = contigs[0][1000:3000]
        for c in contigs[1:]:
            if len(c) > 12000:
                c[5000:7000] = src
    lut = np.zeros(256, np.uint8)
    lut[synth.ACGTN] = np.arange(5)
    reads = []
    for i in range(n_reads):
        if rng.random() < junk_frac:
            reads.append(synth.ACGTN[rng.integers(0, 4, read_len)])
            continue
        ci = int(rng.integers(0, len(contigs)))
        c = contigs[ci]
        u = rng.random()
        if u < 0.02:
            st = int(rng.integers(0, 40))                      # contig start
        elif u < 0.04:
            st = len(c) - read_len - 30 + int(rng.integers(0, 30))  # contig end
        elif u < 0.10 and repeat:
            st = (1000 if ci == 0 else 5000) + int(rng.integers(0, 1800))  # inside the repeat
        else:
            st = int(rng.integers(0, len(c) - read_len - 30))
        codes = lut[c[st:st + read_len + 30]]
        codes = synth.mutate_codes(rng, codes, sub + indel, sub=sub / (sub + indel), dele=indel / (sub + indel) / 2)[:read_len]
        if len(codes) < read_len:
            codes = np.concatenate([codes, rng.integers(0, 4, read_len - len(codes)).astype(np.uint8)])
        if rng.random() < n_frac * 50:
            codes = codes.copy()
            codes[int(rng.integers(0, read_len))] = 4
        if rng.random() < 0.5:
            codes = np.where(codes < 4, 3 - codes, 4)[::-1]
        reads.append(synth.ACGTN[codes])
    return contigs, np.ascontiguousarray(np.stack(reads))


def write_fasta(path, contigs):
    with open(path, "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">chr%d\n" % (i + 1))
            c = np.ascontiguousarray(c, np.uint8)
            rows = len(c) // 80
            if rows:  # 80 bases per line, laid out with numpy (a 3.1 Gbp genome is 39 M lines)
                out = np.empty((rows, 81), np.uint8)
                out[:, :80] = c[:rows * 80].reshape(rows, 80)
                out[:, 80] = 10
                f.write(out.tobytes())
            if len(c) > rows * 80:
                f.write(c[rows * 80:].tobytes() + b"\n")


def write_fastq(path, reads):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            s = bytes(r).decode()
            f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)))


# ------------------------------------------------------------------------------------------------
# the reference program + its call trace
# ------------------------------------------------------------------------------------------------
def have_ref_program():
    return os.path.exists(REF_SR)


def make_long_dataset(seed=1, contig_lens=(1500000, 800000), n_reads=60, read_len=15000, sub=0.005, indel=0.005, sv_frac=0.3,
                      len_jitter=0.2):
    """HiFi / ONT-like reads (ragged lengths) with substitutions and indels; a fraction carries a large deletion or a
    jump (so that the second voting round and the candidate chaining of LR/map.c:1402-1590 are exercised) and some are
    chimeric or unmappable."""
    rng = np.random.default_rng(seed)
    contigs = [synth.random_genome(n, seed=seed * 100 + i) for i, n in enumerate(contig_lens)]
    lut = np.zeros(256, np.uint8)
    lut[synth.ACGTN] = np.arange(5)
    reads = []
    for i in range(n_reads):
        L = int(read_len * (1 + len_jitter * (rng.random() - 0.5)))
        u = rng.random()
        if u < 0.04:
            codes = rng.integers(0, 4, L).astype(np.uint8)
        else:
            ci = int(rng.integers(0, len(contigs)))
            c = contigs[ci]
            span = int(L * 1.15) + 8000
            st = int(rng.integers(0, len(c) - span)) if rng.random() > 0.06 else int(rng.integers(0, 2) * (len(c) - span))
            src = lut[c[st:st + span]]
            if u < 0.04 + sv_frac:
                cut = int(rng.integers(L // 4, 3 * L // 4))
                gap = int(rng.choice([300, 1500, 5000]))
                if rng.random() < 0.5:
                    src = np.concatenate([src[:cut], src[cut + gap:]])           # deletion in the read
                else:
                    src = np.concatenate([src[:cut], rng.integers(0, 4, gap).astype(np.uint8), src[cut:]])  # insertion
            elif u < 0.08 + sv_frac:
                c2 = contigs[int(rng.integers(0, len(contigs)))]
                st2 = int(rng.integers(0, len(c2) - L))
                src = np.concatenate([src[:L // 2], lut[c2[st2:st2 + L]]])       # chimeric
            codes = synth.mutate_codes(rng, src, sub + indel, sub=sub / (sub + indel), dele=indel / (sub + indel) / 2)[:L]
        if rng.random() < 0.5:
            codes = np.where(codes < 4, 3 - codes, 4)[::-1]
        reads.append(synth.ACGTN[codes])
    return contigs, reads


def run_reference(contigs, reads, flags, threads=1, trace=True, workdir=None, program=None):
    """Runs the unmodified reference (GDiet_avx build) on the dataset; returns (sam_text, trace_records)."""
    tmp = workdir or tempfile.mkdtemp(prefix="gdref_")
    fa, fq, sam, tr = (os.path.join(tmp, x) for x in ("ref.fa", "reads.fq", "out.sam", "trace.bin"))
    write_fasta(fa, contigs)
    write_fastq(fq, reads)
    env = dict(os.environ)
    if trace:
        env["GDREF_TRACE"] = tr
    subprocess.run([program or REF_SR, "-t", str(threads)] + flags + ["-o", sam, fa, fq], check=True, env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(sam) as f:
        sam_text = f.read()
    return sam_text, (parse_trace(tr) if trace else None)


def parse_trace(path):
    """-> list (one per read, input order) of dicts {seq, cands: [ {exact, q, t, w, score, cigar, rid, qs, qe, rs, re, rev} ]}"""
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    reads = []

    def i32(n=1):
        nonlocal pos
        v = struct.unpack_from("<%di" % n, buf, pos)
        pos += 4 * n
        return v if n > 1 else v[0]

    cur = None
    while pos < len(buf):
        kind = i32()
        if kind == 1:
            n = i32()
            reads.append(dict(seq=buf[pos:pos + n], cands=[]))
            pos += n
            cur = None
        elif kind == 2:
            n = i32()
            q = np.frombuffer(buf, np.uint8, n, pos)
            t = np.frombuffer(buf, np.uint8, n, pos + n)
            pos += 2 * n
            ex = i32()
            cur = dict(exact=ex, q=q, t=t, w=None, score=None, cigar=None)
            if ex:
                reads[-1]["cands"].append(cur)
        elif kind == 3:
            qlen, tlen, w, zdrop, end_bonus, flag, gq, ge, gq2, ge2 = i32(10)
            q = np.frombuffer(buf, np.uint8, qlen, pos)
            t = np.frombuffer(buf, np.uint8, tlen, pos + qlen)
            pos += qlen + tlen
            score, nc = i32(2)
            cig = np.frombuffer(buf, np.uint32, max(nc, 0), pos).copy()
            pos += 4 * max(nc, 0)
            cur = dict(exact=0, q=q, t=t, w=w, score=score, cigar=cig, flag=flag, zdrop=zdrop, end_bonus=end_bonus)
            reads[-1]["cands"].append(cur)
        elif kind == 4:
            rid, score, qs, qe, rs, re_, rev = i32(7)
            c = reads[-1]["cands"][-1]
            c.update(rid=rid, qs=qs, qe=qe, rs=rs, re=re_, rev=rev, has_extra=True)
            if c["exact"]:
                c["score"] = score
                c["cigar"] = np.array([len(c["q"]) << 4], np.uint32)
        else:
            raise ValueError("bad trace record %d at %d" % (kind, pos))
    return reads


# ------------------------------------------------------------------------------------------------
# oracle binding (oracle/gd_oracle_map.c)
# ------------------------------------------------------------------------------------------------
class MapOracle:
    def __init__(self):
        self.lib = L = Oracle().lib
        L.gdo_index_build.restype = C.c_void_p
        L.gdo_index_build.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.gdo_index_destroy.argtypes = [C.c_void_p]
        L.gdo_index_get.restype = C.POINTER(C.c_uint64)
        L.gdo_index_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        L.gdo_index_cal_max_occ.restype = C.c_int32
        L.gdo_index_cal_max_occ.argtypes = [C.c_void_p, C.c_float]
        L.gdo_lr_map_read.restype = C.c_int
        L.gdo_lr_map_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(gd.gd_lr_opt_t), C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p]
        L.gdo_sr_map_read.restype = C.c_int
        L.gdo_sr_map_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(SrOpt), C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p]

    def index_build(self, contigs, w, k, Z):
        buf = b"".join(c.tobytes() for c in contigs)
        lens = np.array([len(c) for c in contigs], np.int32)
        off = np.zeros(len(contigs), np.int64)
        off[1:] = np.cumsum(lens[:-1])
        self._keep = (buf, lens, off)
        return self.lib.gdo_index_build(len(contigs), buf, off.ctypes.data, lens.ctypes.data, w, k, Z.encode(), len(Z))

    def index_get(self, mi, minier):
        n = C.c_int(0)
        p = self.lib.gdo_index_get(mi, int(minier), C.byref(n))
        return np.array([p[i] for i in range(n.value)], np.uint64)

    def index_arrays(self, mi):
        """(keys, counts, positions) of the oracle index."""
        class Idx(C.Structure):
            _fields_ = [("n_seq", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("len", C.c_void_p), ("offset", C.c_void_p),
                        ("codes", C.c_void_p), ("n_keys", C.c_int64), ("key", C.POINTER(C.c_uint64)),
                        ("start", C.POINTER(C.c_uint64)), ("cnt", C.POINTER(C.c_uint32)), ("n_pos", C.c_int64),
                        ("pos", C.POINTER(C.c_uint64))]
        s = C.cast(mi, C.POINTER(Idx)).contents
        nk, npos = s.n_keys, s.n_pos
        return (np.ctypeslib.as_array(s.key, (nk,)).copy(), np.ctypeslib.as_array(s.cnt, (nk,)).copy(),
                np.ctypeslib.as_array(s.pos, (npos,)).copy())

    def lr_map_read(self, mi, seq, opt, cap=16):
        out = np.zeros(cap, CAND_DTYPE)
        cig = np.zeros(cap * (2 * len(seq) + 4096), np.uint32)
        dbg = np.zeros(1, DBG_DTYPE)
        n = self.lib.gdo_lr_map_read(mi, bytes(seq), len(seq), C.byref(opt), out.ctypes.data, cap, cig.ctypes.data, len(cig),
                                     dbg.ctypes.data)
        return out[:n], cig, dbg[0]

    def map_read(self, mi, seq, opt, cap=64):
        out = np.zeros(cap, CAND_DTYPE)
        cig = np.zeros(cap * (2 * len(seq) + 8), np.uint32)
        dbg = np.zeros(1, DBG_DTYPE)
        n = self.lib.gdo_sr_map_read(mi, bytes(seq), len(seq), C.byref(opt), out.ctypes.data, cap, cig.ctypes.data, len(cig),
                                     dbg.ctypes.data)
        return out[:n], cig, dbg[0]


def lr_cands_equal_trace(cands, cig, tr_cands, what=""):
    """Long reads: every DP call of the reference in order (lengths, score, CIGAR) and, where the reference went on to
    mm_update_extra (score != KSW_NEG_INF), the candidate window."""
    assert len(cands) == len(tr_cands), "%s: %d candidates, reference made %d DP calls" % (what, len(cands), len(tr_cands))
    for j, (c, t) in enumerate(zip(cands, tr_cands)):
        got = (int(c["qe"] - c["qs"]), int(c["re"] - c["rs"]), int(c["score"]), int(c["exact"]))
        exp = (len(t["q"]), len(t["t"]), t["score"], t["exact"])
        assert got == exp, "%s cand %d: got (qlen, tlen, score, exact) %s, reference %s" % (what, j, got, exp)
        if t.get("has_extra"):
            g2 = tuple(int(c[f]) for f in ("rid", "qs", "qe", "rs", "re", "rev"))
            e2 = (t["rid"], t["qs"], t["qe"], t["rs"], t["re"], t["rev"])
            assert g2 == e2, "%s cand %d: window %s, reference %s" % (what, j, g2, e2)
        gc = cig[int(c["cigar_off"]):int(c["cigar_off"]) + max(int(c["n_cigar"]), 0)]
        assert np.array_equal(gc, t["cigar"]), "%s cand %d cigar differs (%d vs %d entries)" % (what, j, len(gc), len(t["cigar"]))


def cands_equal_trace(cands, cig, tr_cands, what=""):
    """Compare oracle/device candidates of one read with the reference trace of that read."""
    assert len(cands) == len(tr_cands), "%s: %d candidates, reference made %d" % (what, len(cands), len(tr_cands))
    for j, (c, t) in enumerate(zip(cands, tr_cands)):
        got = tuple(int(c[f]) for f in ("rid", "qs", "qe", "rs", "re", "rev", "exact", "score"))
        exp = (t["rid"], t["qs"], t["qe"], t["rs"], t["re"], t["rev"], t["exact"], t["score"])
        assert got == exp, "%s cand %d: got %s, reference %s" % (what, j, got, exp)
        gc = cig[int(c["cigar_off"]):int(c["cigar_off"]) + max(int(c["n_cigar"]), 0)]
        assert np.array_equal(gc, t["cigar"]), "%s cand %d cigar: got %s, reference %s" % (what, j, gc, t["cigar"])
