"""ctypes loaders for the CHECKERS (tests / smoke / bench cpu_baseline only -- never the product).

* ``Oracle``  -> oracle/libgd_oracle.so   our scalar restatement (oracle/gd_oracle.c)
* ``Ref``     -> oracle/_ref/libgdref_{avx,scalar}.so   the unmodified reference, compiled by
                 oracle/Makefile from /root/reference (only in the build container; the .so travels).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

EXTZ_FIELDS = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end"]
EXTZ_DTYPE = np.dtype([(f, np.int32) for f in EXTZ_FIELDS])

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return "avx512bw" in txt and "avx512dq" in txt
    except OSError:
        return False


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def build_ref():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


def sr_mat(a=2, b=8):
    """5x5 matrix built like GDiet-ShortReads/map.c:861-865."""
    bb = -abs(b)
    m = np.full((5, 5), bb, np.int8)
    np.fill_diagonal(m, a)
    m[4, :] = 0
    m[:, 4] = 0
    return m.reshape(-1).copy()


class _DPMixin:
    def _extd2(self, fn, lead, q, t, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, tail=()):
        q = np.ascontiguousarray(q, np.uint8)
        t = np.ascontiguousarray(t, np.uint8)
        mat = np.ascontiguousarray(mat, np.int8)
        m = int(round(len(mat) ** 0.5))
        ez = np.zeros(1, EXTZ_DTYPE)
        cap = len(q) + len(t) + 8
        cig = np.zeros(cap, np.uint32)
        qq = q if len(q) else np.zeros(1, np.uint8)
        tt = t if len(t) else np.zeros(1, np.uint8)
        n = fn(*lead, len(q), qq, len(t), tt, m, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, *tail,
               ez.ctypes.data_as(C.c_void_p), cig, cap)
        assert n >= 0, "cigar overflow"
        return {f: int(ez[0][f]) for f in EXTZ_FIELDS}, cig[: max(n, 0)].copy()


class Oracle(_DPMixin):
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libgd_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.lib = C.CDLL(path)
        L.gdo_ksw_extd2.restype = C.c_int
        L.gdo_ksw_extd2.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, i8p] + [C.c_int] * 9 + [C.c_void_p, u32p, C.c_int]
        L.gdo_band_cells.restype = C.c_int64
        L.gdo_band_cells.argtypes = [C.c_int] * 3
        L.gdo_exact_match.restype = C.c_int
        L.gdo_exact_match.argtypes = [C.c_int, u8p, C.c_int, u8p]
        L.gdo_mm_sketch.restype = C.c_long
        L.gdo_mm_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_int, u64p, C.c_long]
        L.gdo_mm_sketch3.restype = C.c_long
        L.gdo_mm_sketch3.argtypes = [C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_int, C.c_int,
                                     C.c_uint32, u64p, C.c_long, C.POINTER(C.c_uint32)]
        L.gdo_mm_sketch2.restype = C.c_long
        L.gdo_mm_sketch2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_int, C.c_float,
                                     u64p, C.c_long, u32p]
        L.gdo_hash64.restype = C.c_uint64
        L.gdo_hash64.argtypes = [C.c_uint64, C.c_uint64]

    def ksw_extd2(self, q, t, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, score_rule=1):
        return self._extd2(self.lib.gdo_ksw_extd2, (), q, t, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag,
                           tail=(score_rule,))

    def band_cells(self, qlen, tlen, w):
        return int(self.lib.gdo_band_cells(qlen, tlen, w))

    def exact_match(self, q, t):
        return int(self.lib.gdo_exact_match(len(q), np.ascontiguousarray(q, np.uint8), len(t),
                                            np.ascontiguousarray(t, np.uint8)))

    def mm_sketch(self, seq, w, k, rid, Z):
        return _sketch1(self.lib.gdo_mm_sketch, seq, w, k, rid, Z)

    def mm_sketch3(self, seq, w, k, rid, Z, shift, max_nb_seeds):
        return _sketch3(self.lib.gdo_mm_sketch3, seq, w, k, rid, Z, shift, max_nb_seeds)

    def mm_sketch2(self, seq, w, k, rid, Z, max_seeds):
        return _sketch2(self.lib.gdo_mm_sketch2, seq, w, k, rid, Z, max_seeds)


def _as_bytes(seq):
    return seq if isinstance(seq, bytes) else bytes(seq)


def _sketch1(fn, seq, w, k, rid, Z):
    seq = _as_bytes(seq)
    Zb = Z.encode() if isinstance(Z, str) else Z
    cap = len(seq) + 8
    out = np.zeros(2 * cap, np.uint64)
    n = fn(seq, len(seq), w, k, rid, Zb, len(Zb), out, cap)
    assert n <= cap
    return out[: 2 * n].reshape(-1, 2).copy()


def _sketch3(fn, seq, w, k, rid, Z, shift, max_nb_seeds):
    seq = _as_bytes(seq)
    Zb = Z.encode() if isinstance(Z, str) else Z
    cap = len(seq) + 8
    out = np.zeros(2 * cap, np.uint64)
    ret = C.c_uint32(0)
    n = fn(seq, len(seq), w, k, rid, Zb, len(Zb), shift, max_nb_seeds, out, cap, C.byref(ret))
    assert n <= cap
    return out[: 2 * n].reshape(-1, 2).copy(), int(ret.value)


def _sketch2(fn, seq, w, k, rid, Z, max_seeds):
    seq = _as_bytes(seq)
    Zb = Z.encode() if isinstance(Z, str) else Z
    cap = len(seq) * len(Zb) + 8
    out = np.zeros(2 * cap, np.uint64)
    counts = np.zeros(len(Zb), np.uint32)
    n = fn(seq, len(seq), w, k, rid, Zb, len(Zb), max_seeds, out, cap, counts)
    assert n <= cap
    return out[: 2 * n].reshape(-1, 2).copy(), counts


class Ref(_DPMixin):
    """The unmodified reference. variant: 'avx' (GDiet_avx objects) or 'scalar'."""

    def __init__(self, variant="avx"):
        path = os.path.join(ORACLE_DIR, "_ref", "libgdref_%s.so" % variant)
        if not os.path.exists(path):
            if os.path.isdir("/root/reference"):
                build_ref()
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        if variant == "avx" and not cpu_has_avx512():
            raise RuntimeError("host CPU lacks AVX-512; use Ref('scalar')")
        self.variant = variant
        L = self.lib = C.CDLL(path)
        L.ref_ksw_extd2.restype = C.c_int
        L.ref_ksw_extd2.argtypes = [C.c_int, C.c_void_p, C.c_int, u8p, C.c_int, u8p, C.c_int, i8p] + [C.c_int] * 8 + [
            C.c_void_p, u32p, C.c_int]
        L.ref_exact_match.restype = C.c_int
        L.ref_exact_match.argtypes = [C.c_int, u8p, C.c_int, u8p]
        L.ref_ksw_extd2_batch.restype = C.c_int
        L.ref_ksw_extd2_batch.argtypes = [C.c_int, C.c_int, i32p, i64p, u8p, i32p, i64p, u8p, C.c_int, i8p] + [C.c_int] * 8 + [
            C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_mm_sketch.restype = C.c_long
        L.ref_mm_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_int, u64p, C.c_long]
        L.ref_mm_sketch3.restype = C.c_long
        L.ref_mm_sketch3.argtypes = [C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_int, C.c_int,
                                     C.c_uint32, u64p, C.c_long, C.POINTER(C.c_uint32)]
        L.ref_mm_sketch2.restype = C.c_long
        L.ref_mm_sketch2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_int, C.c_float,
                                     u64p, C.c_long, u32p]
        L.ref_mm_sketch_batch.restype = C.c_int
        L.ref_mm_sketch_batch.argtypes = [C.c_int, i64p, i32p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int, i64p, C.c_int]

    def ksw_extd2(self, q, t, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, which=None):
        if which is None:
            which = 1 if self.variant == "avx" else 0
        return self._extd2(self.lib.ref_ksw_extd2, (which, None), q, t, mat, gapo, gape, gapo2, gape2, w, zdrop,
                           end_bonus, flag)

    def ksw_extd2_batch(self, qlen, qoff, qbuf, tlen, toff, tbuf, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus,
                        flag, n_threads, cigar_stride=0, which=None, out=None):
        """out = (ez, cig) preallocated by the caller (a timed loop must not pay for a fresh 1.4 GB array per step)"""
        if which is None:
            which = 1 if self.variant == "avx" else 0
        n = len(qlen)
        if out is not None:
            ez, cig = out
            assert len(ez) >= n and (not cigar_stride or len(cig) >= n * cigar_stride)
        else:
            ez = np.zeros(n, EXTZ_DTYPE)
            cig = np.zeros(n * cigar_stride, np.uint32) if cigar_stride else None
        m = int(round(len(mat) ** 0.5))
        self.lib.ref_ksw_extd2_batch(which, n, qlen, qoff, qbuf, tlen, toff, tbuf, m, mat, gapo, gape, gapo2, gape2, w,
                                     zdrop, end_bonus, flag, ez.ctypes.data_as(C.c_void_p),
                                     cig.ctypes.data_as(C.c_void_p) if cig is not None else None, cigar_stride,
                                     n_threads)
        return ez, cig

    def exact_match(self, q, t):
        return int(self.lib.ref_exact_match(len(q), np.ascontiguousarray(q, np.uint8), len(t),
                                            np.ascontiguousarray(t, np.uint8)))

    def mm_sketch(self, seq, w, k, rid, Z):
        return _sketch1(self.lib.ref_mm_sketch, seq, w, k, rid, Z)

    def mm_sketch3(self, seq, w, k, rid, Z, shift, max_nb_seeds):
        return _sketch3(self.lib.ref_mm_sketch3, seq, w, k, rid, Z, shift, max_nb_seeds)

    def mm_sketch2(self, seq, w, k, rid, Z, max_seeds):
        return _sketch2(self.lib.ref_mm_sketch2, seq, w, k, rid, Z, max_seeds)

    def mm_sketch_batch(self, off, lens, buf, w, k, Z, n_threads):
        Zb = Z.encode() if isinstance(Z, str) else Z
        counts = np.zeros(len(lens), np.int64)
        self.lib.ref_mm_sketch_batch(len(lens), off, lens, buf, w, k, Zb, len(Zb), counts, n_threads)
        return counts
