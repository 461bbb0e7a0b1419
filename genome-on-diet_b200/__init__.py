"""genome-on-diet_b200 -- host-side Python mirror of the C ABI in include/gdiet_cuda.h.

The product is libgdiet_cuda.so (hand-written sm_100a CUDA behind a C ABI that keeps the reference's
``ksw_extd2_sse()/ksw_extz_t`` and ``mm_sketch*()/mm128_v`` signatures). This module only binds it with
ctypes so that tests and bench.py can call exactly what a C host program would call. There is no
CPU fallback anywhere: if the shared library is missing or no sm_100 GPU is usable, every entry
point raises.

Import it as ``import gdiet_b200`` (repo-root shim; the directory name contains a hyphen).
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libgdiet_cuda.so")
CSRC_DIR = os.path.join(PKG_DIR, "csrc")

# KSW_EZ_* (GDiet-ShortReads/ksw2.h:9-18)
KSW_EZ_SCORE_ONLY = 0x01
KSW_EZ_RIGHT = 0x02
KSW_EZ_GENERIC_SC = 0x04
KSW_EZ_APPROX_MAX = 0x08
KSW_EZ_APPROX_DROP = 0x10
KSW_EZ_EXTZ_ONLY = 0x40
KSW_EZ_REV_CIGAR = 0x80
KSW_NEG_INF = -0x40000000

GD_OK, GD_ERR_NO_DEVICE, GD_ERR_CUDA, GD_ERR_ARG, GD_ERR_CAPACITY = 0, 1, 2, 3, 4

EXTZ_FIELDS = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar", "reach_end"]
GD_EXTZ_DTYPE = np.dtype([(f, np.int32) for f in EXTZ_FIELDS + ["tb_i", "tb_j", "rows_done", "lead64", "r1"]])
MM128_DTYPE = np.dtype([("x", np.uint64), ("y", np.uint64)])


class GdietError(RuntimeError):
    pass


class ksw_extz_t(C.Structure):
    """GDiet-ShortReads/ksw2.h:31-40"""
    _fields_ = [("max_zdropped", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int),
                ("mqe_t", C.c_int), ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int), ("m_cigar", C.c_int),
                ("n_cigar", C.c_int), ("reach_end", C.c_int), ("cigar", C.POINTER(C.c_uint32))]


class mm128_v(C.Structure):
    """GDiet-ShortReads/minimap.h:72-76"""
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


class mm_pattern_t(C.Structure):
    """GDiet-ShortReads/minimap.h:99-102"""
    _fields_ = [("n", C.c_uint32), ("shift_seeds_number", C.POINTER(C.c_uint32))]


class gd_ksw_params_t(C.Structure):
    _fields_ = [("m", C.c_int32), ("mat", C.c_void_p), ("q", C.c_int32), ("e", C.c_int32), ("q2", C.c_int32),
                ("e2", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("flag", C.c_int32)]


class gd_sr_opt_t(C.Structure):
    """include/gdiet_cuda.h: the mm_mapopt_t fields the short-read path reads (GDiet-ShortReads/minimap.h:142-205)."""
    _fields_ = [("W", C.c_int32), ("Z", C.c_char * 64), ("max_seeds", C.c_float), ("frag_mode", C.c_int32),
                ("max_frag_len", C.c_int32), ("bw", C.c_uint32), ("min_cnt", C.c_float), ("rec_threshold_frac", C.c_float),
                ("af_max_loc", C.c_int32), ("mid_occ", C.c_int32), ("max_max_occ", C.c_int32), ("occ_dist", C.c_int32),
                ("q_occ_frac", C.c_float), ("for_only", C.c_int32), ("rev_only", C.c_int32),
                ("a", C.c_int32), ("b", C.c_int32), ("q", C.c_int32), ("e", C.c_int32), ("q2", C.c_int32), ("e2", C.c_int32),
                ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("bw_frac", C.c_float), ("bw_min", C.c_uint32), ("bw_max", C.c_uint32)]


GD_INDEX_NBUF = 7


class gd_index_meta_t(C.Structure):
    _fields_ = [(f, C.c_int64) for f in ("n_seq", "total_len", "n_minimizers", "n_keys", "table_slots", "s_words")] + [
        ("w", C.c_int32), ("k", C.c_int32)]


class gd_sr_post_opt_t(C.Structure):
    """include/gdiet_cuda.h: options of the host side after the DP (GDiet-ShortReads/map.c:954-984)."""
    _fields_ = [(f, C.c_int32) for f in ("a", "b", "q", "e", "min_dp_max", "best_n", "no_print_2nd", "is_sr", "sam_hit_only",
                                         "softclip", "n_threads", "q2", "e2")]


def sr_post_options(n_threads=0, **kw):
    """`-ax sr` values (GDiet-ShortReads/options.c:130-150)."""
    o = gd_sr_post_opt_t(a=2, b=8, q=12, e=2, min_dp_max=40, best_n=20, no_print_2nd=1, is_sr=1, sam_hit_only=0, softclip=0,
                         n_threads=n_threads, q2=24, e2=1)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _cstr_array(strings):
    if isinstance(strings, C.Array):
        return strings
    arr = (C.c_char_p * len(strings))()
    arr[:] = [x if isinstance(x, bytes) else x.encode() for x in strings]
    return arr


class SamText:
    """The malloc'ed SAM text gd_sr_sam_batch returns (a C host writes it out and calls gd_free)."""

    def __init__(self, ptr, n):
        self.ptr, self.n = ptr, n

    def bytes(self):
        return C.string_at(self.ptr, self.n)

    def free(self):
        if self.ptr:
            load().gd_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.free()


def lr_post_options(preset="map-hifi", n_threads=0, **kw):
    """`-ax map-hifi | map-ont` values (GDiet-LongReads/options.c:86-111).  min_dp_max is 40 for every preset: the
    long-read main() assigns it AFTER the option pass that applies -x (LR/main.c:134-160,181), so the preset's 200 is
    overwritten; only -s changes it."""
    sc = dict(a=1, b=4, q=6, e=2, q2=26, e2=1, min_dp_max=40) if preset == "map-hifi" else dict(a=2, b=4, q=4, e=2, q2=24, e2=1, min_dp_max=40)
    o = gd_sr_post_opt_t(best_n=5, no_print_2nd=0, is_sr=0, sam_hit_only=0, softclip=0, n_threads=n_threads, **sc)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def flat_ref(contigs):
    """(ref_off, ref_len, one concatenated buffer) of a list of contigs: what a C host already holds; pass it as ref= to the
    *_sam_batch wrappers so that a timed call does not concatenate gigabytes of reference again"""
    ref_len = np.array([len(c) for c in contigs], np.int32)
    ref_off = np.zeros(len(contigs), np.int64)
    ref_off[1:] = np.cumsum(ref_len[:-1].astype(np.int64))
    ref = np.concatenate([np.ascontiguousarray(c, np.uint8) for c in contigs]) if len(contigs) > 1 else np.ascontiguousarray(contigs[0], np.uint8)
    return ref_off, ref_len, ref


def lr_sam_batch(names, off, lens, seq, qual, cand_off, cand, cigar, seq_names, contigs, opt, ref=None):
    """gd_lr_sam_batch. Returns (SAM bytes, sam_off[n+1], needs_stitch[n])."""
    L = load()
    n = len(lens)
    ref_off, ref_len, ref = ref if ref is not None else flat_ref(contigs)
    n_arr, s_arr = _cstr_array(names), _cstr_array(seq_names)
    out, out_len = C.c_void_p(), C.c_size_t(0)
    cand = np.ascontiguousarray(cand) if len(cand) else np.zeros(1, SR_CAND_DTYPE)
    cigar = np.ascontiguousarray(cigar, np.uint32) if len(cigar) else np.zeros(1, np.uint32)
    sam_off, stitch = np.zeros(n + 1, np.int64), np.zeros(max(n, 1), np.uint8)
    rc = L.gd_lr_sam_batch(n, C.cast(n_arr, C.c_void_p), _ptr(off), _ptr(lens), _ptr(seq), _ptr(qual), _ptr(cand_off), _ptr(cand),
                           _ptr(cigar), len(ref_len), C.cast(s_arr, C.c_void_p), _ptr(ref_off), _ptr(ref_len), _ptr(ref), C.byref(opt),
                           C.byref(out), C.byref(out_len), _ptr(sam_off), _ptr(stitch))
    if rc != GD_OK:
        raise GdietError("gd_lr_sam_batch failed (%d)" % rc)
    txt = C.string_at(out, out_len.value)
    L.gd_free(out)
    return txt, sam_off, stitch[:n]


class SamParts:
    """The pieces gd_sr_sam_batch_parts returns (input order); bytes() joins them, free() releases them."""

    def __init__(self, parts, lens, n):
        self.parts, self.lens, self.count = parts, lens, n
        self.n = sum(int(lens[i]) for i in range(n))

    def bytes(self):
        return b"".join(C.string_at(self.parts[i], self.lens[i]) for i in range(self.count))

    def free(self):
        if self.parts:
            L = load()
            for i in range(self.count):
                L.gd_free(self.parts[i])
            L.gd_free(C.cast(self.parts, C.c_void_p)), L.gd_free(C.cast(self.lens, C.c_void_p))
            self.parts = None

    def __del__(self):
        self.free()


def sr_sam_batch(names, off, lens, seq, qual, cand_off, cand, cigar, seq_names, contigs, opt, raw=False, ref=None, parts=False):
    """gd_sr_sam_batch: SAM records of a mapped batch (bytes, or a SamText handle with raw=True).
    contigs: list of ASCII uint8 arrays; ref = (ref_off, ref_len, concatenated buffer) may be passed to reuse it."""
    L = load()
    if ref is not None:
        ref_off, ref_len, ref = ref
    else:
        ref_len = np.array([len(c) for c in contigs], np.int32)
        ref_off = np.zeros(len(contigs), np.int64)
        ref_off[1:] = np.cumsum(ref_len[:-1].astype(np.int64))
        ref = np.concatenate([np.ascontiguousarray(c, np.uint8) for c in contigs]) if len(contigs) > 1 else np.ascontiguousarray(contigs[0], np.uint8)
    n_arr, s_arr = _cstr_array(names), _cstr_array(seq_names)
    out, out_len = C.c_void_p(), C.c_size_t(0)
    cand = np.ascontiguousarray(cand)
    cigar = np.ascontiguousarray(cigar, np.uint32) if len(cigar) else np.zeros(1, np.uint32)
    if len(cand) == 0:
        cand = np.zeros(1, SR_CAND_DTYPE)
    if parts:
        pp, pl, pn = C.POINTER(C.c_void_p)(), C.POINTER(C.c_size_t)(), C.c_int32(0)
        rc = L.gd_sr_sam_batch_parts(len(lens), C.cast(n_arr, C.c_void_p), _ptr(off), _ptr(lens), _ptr(seq), _ptr(qual), _ptr(cand_off),
                                     _ptr(cand), _ptr(cigar), len(ref_len), C.cast(s_arr, C.c_void_p), _ptr(ref_off), _ptr(ref_len),
                                     _ptr(ref), C.byref(opt), C.byref(pp), C.byref(pl), C.byref(pn))
        if rc != GD_OK:
            raise GdietError("gd_sr_sam_batch_parts failed (%d)" % rc)
        return SamParts(pp, pl, pn.value)
    rc = L.gd_sr_sam_batch(len(lens), C.cast(n_arr, C.c_void_p), _ptr(off), _ptr(lens), _ptr(seq), _ptr(qual), _ptr(cand_off),
                           _ptr(cand), _ptr(cigar), len(ref_len), C.cast(s_arr, C.c_void_p), _ptr(ref_off), _ptr(ref_len), _ptr(ref),
                           C.byref(opt), C.byref(out), C.byref(out_len))
    if rc != GD_OK:
        raise GdietError("gd_sr_sam_batch failed (%d)" % rc)
    h = SamText(out, out_len.value)
    if raw:
        return h
    txt = h.bytes()  # (one more copy, into a Python bytes object)
    h.free()
    return txt


def mmi_write(path, w, k, names, lens, keys, counts, positions, S, bucket_bits=14, flag=0):
    """gd_mmi_write: the reference's .mmi file from the exported index arrays."""
    L = load()
    n_arr = _cstr_array(names)
    lens = np.ascontiguousarray(lens, np.int32)
    rc = L.gd_mmi_write(path.encode(), w, k, bucket_bits, flag, len(names), C.cast(n_arr, C.c_void_p), _ptr(lens), len(keys),
                        _ptr(np.ascontiguousarray(keys, np.uint64)), _ptr(np.ascontiguousarray(counts, np.uint32)),
                        _ptr(np.ascontiguousarray(positions, np.uint64)), _ptr(np.ascontiguousarray(S, np.uint32)))
    if rc != GD_OK:
        raise GdietError("gd_mmi_write failed (%d)" % rc)


def sam_header(seq_names, ref_len):
    L = load()
    s_arr = _cstr_array(seq_names)
    ref_len = np.ascontiguousarray(ref_len, np.int32)
    out, out_len = C.c_void_p(), C.c_size_t(0)
    rc = L.gd_sam_header(len(seq_names), C.cast(s_arr, C.c_void_p), _ptr(ref_len), C.byref(out), C.byref(out_len))
    if rc != GD_OK:
        raise GdietError("gd_sam_header failed (%d)" % rc)
    txt = C.string_at(out, out_len.value)
    L.gd_free(out)
    return txt


class gd_lr_opt_t(C.Structure):
    """include/gdiet_cuda.h: the mm_mapopt_t fields the long-read path reads (GDiet-LongReads/map.c:1273-1853)."""
    _fields_ = [("W", C.c_int32), ("Z", C.c_char * 64), ("max_seeds", C.c_float), ("frag_mode", C.c_int32),
                ("max_frag_len", C.c_int32), ("bw", C.c_uint32), ("mid_occ", C.c_int32), ("max_max_occ", C.c_int32),
                ("occ_dist", C.c_int32), ("q_occ_frac", C.c_float), ("for_only", C.c_int32), ("rev_only", C.c_int32),
                ("a", C.c_int32), ("b", C.c_int32), ("q", C.c_int32), ("e", C.c_int32), ("q2", C.c_int32), ("e2", C.c_int32),
                ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("vt_dis", C.c_uint32), ("vt_nb_loc", C.c_uint32),
                ("vt_cov", C.c_float), ("vt_df1", C.c_float), ("vt_df2", C.c_float), ("vt_f", C.c_float),
                ("max_max_gap", C.c_uint32), ("max_min_gap", C.c_uint32)]


def lr_options(preset="map-hifi", Z="10", bw=1000, mid_occ=50, **kw):
    """`-ax map-hifi|map-ont -Z .. -W .. -r bw` (GDiet-LongReads/options.c:86-111, main.c:170-182). mid_occ is what
    mm_mapopt_update derives from the index (gd_index_cal_max_occ clamped to [min_mid_occ, max_mid_occ])."""
    sc = dict(a=1, b=4, q=6, e=2, q2=26, e2=1) if preset == "map-hifi" else dict(a=2, b=4, q=4, e=2, q2=24, e2=1)
    o = gd_lr_opt_t(W=len(Z), Z=Z.encode(), max_seeds=0.1, frag_mode=0, max_frag_len=0, bw=bw, mid_occ=mid_occ, max_max_occ=4095,
                    occ_dist=500, q_occ_frac=0.01, for_only=0, rev_only=0, zdrop=400, end_bonus=-1, vt_dis=100, vt_nb_loc=3,
                    vt_cov=0.03, vt_df1=0.01, vt_df2=0.01, vt_f=0.05, max_max_gap=50000, max_min_gap=4000, **sc)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


SR_CAND_FIELDS = ["rid", "rs", "re", "qs", "qe", "rev", "votes", "first_q", "last_q", "exact", "score", "n_cigar", "cigar_off"]
SR_CAND_DTYPE = np.dtype([(f, np.int32) for f in SR_CAND_FIELDS] + [("reserved", np.int32, 3)])


def sr_options(Z="10", qlen=150, bw_frac=0.05, bw_min=150, bw_max=200, min_cnt=2.0, rec_frac=0.0, af_max_loc=20, **kw):
    """What `-ax sr -Z .. -W .. -r bw_frac,bw_min,bw_max -n min_cnt,rec_frac` leaves in mm_mapopt_t
    (GDiet-ShortReads/options.c:130-150, main.c:166-182).  The band is computed per read from bw_frac, bw_min, bw_max as at
    map.c:624-631 (`bw`, the value for a read of `qlen` bases, is only used when bw_max is set to 0)."""
    bw = int(np.float32(qlen * np.float32(bw_frac)))
    if bw_min > bw:
        bw = bw_min
    elif bw_max < bw:
        bw = bw_max
    o = gd_sr_opt_t(W=len(Z), Z=Z.encode(), max_seeds=0.1, frag_mode=1, max_frag_len=800, bw=bw, min_cnt=min_cnt,
                    rec_threshold_frac=rec_frac, af_max_loc=af_max_loc, mid_occ=1000, max_max_occ=4095, occ_dist=500,
                    q_occ_frac=0.01, for_only=0, rev_only=0, a=2, b=8, q=12, e=2, q2=24, e2=1, zdrop=100, end_bonus=10,
                    bw_frac=bw_frac, bw_min=bw_min, bw_max=bw_max)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build(verbose=False):
    """Compile libgdiet_cuda.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC_DIR, "-j4"], capture_output=True, text=True)
    if out.returncode != 0:
        raise GdietError("building libgdiet_cuda.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None

# every symbol include/gdiet_cuda.h declares
EXPORTS = ["gd_init", "gd_destroy", "gd_strerror", "gd_set_option", "gd_get_stat", "gd_stream", "gd_thread_ctx_pool_size", "ksw_extd2_sse",
           "ksw_extd2_avx512", "gd_ksw_extd2_batch", "gd_ksw_extd2_batch_device", "gd_exact_match_batch_device",
           "mm_sketch", "mm_sketch2", "mm_sketch3", "gd_sketch_ref_batch", "gd_sketch_ref_batch_device",
           "gd_sketch_reads_batch", "gd_index_build", "gd_index_destroy", "gd_index_stat", "gd_index_get_batch",
           "gd_index_export", "gd_index_cal_max_occ", "gd_sr_map_batch", "gd_sr_sam_batch", "gd_sam_header", "gd_free", "gd_index_meta", "gd_index_alloc", "gd_index_buffers", "gd_index_commit", "gd_index_build_device", "gd_lr_map_batch", "gd_mmi_write", "gd_lr_sam_batch", "gd_index_load_mmi", "gd_index_seq_name", "gd_sr_sam_batch_parts",
           "gd_pinned_alloc", "gd_pinned_free", "gd_multi_init", "gd_multi_destroy", "gd_multi_size", "gd_multi_ctx", "gd_multi_index",
           "gd_multi_strerror", "gd_multi_index_bcast", "gd_multi_stat", "gd_multi_prepare_sam", "gd_sr_map_sam_prepare", "gd_multi_sr_map_batch", "gd_multi_lr_map_batch",
           "gd_multi_sr_map_sam", "gd_multi_lr_map_sam", "gd_sr_map_sam_batch"]


def load():
    """dlopen the product library (RTLD_LOCAL). Raises if it was not built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdietError("libgdiet_cuda.so not found at %s -- run __graft_entry__.build() / make -C %s" % (LIB_PATH, CSRC_DIR))
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.gd_init.restype = i32
    L.gd_init.argtypes = [i32, C.POINTER(vp)]
    L.gd_destroy.restype = None
    L.gd_destroy.argtypes = [vp]
    L.gd_strerror.restype = C.c_char_p
    L.gd_strerror.argtypes = [vp]
    L.gd_set_option.restype = i32
    L.gd_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.gd_get_stat.restype = C.c_long
    L.gd_get_stat.argtypes = [vp, C.c_char_p]
    L.gd_stream.restype = vp
    L.gd_stream.argtypes = [vp]
    L.gd_thread_ctx_pool_size.restype = C.c_long
    L.gd_thread_ctx_pool_size.argtypes = [C.c_int]
    ksw_args = [vp, i32, vp, i32, vp, C.c_int8, vp, C.c_int8, C.c_int8, C.c_int8, C.c_int8, i32, i32, i32, i32,
                C.POINTER(ksw_extz_t)]
    for name in ("ksw_extd2_sse", "ksw_extd2_avx512"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = ksw_args
    L.gd_ksw_extd2_batch.restype = i32
    L.gd_ksw_extd2_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(gd_ksw_params_t), vp, vp, vp, i64]
    L.gd_ksw_extd2_batch_device.restype = i32
    L.gd_ksw_extd2_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32,
                                            C.POINTER(gd_ksw_params_t), vp, vp, i32]
    L.gd_exact_match_batch_device.restype = i32
    L.gd_exact_match_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.mm_sketch.restype = None
    L.mm_sketch.argtypes = [vp, C.c_char_p, i32, i32, i32, C.c_uint32, i32, C.POINTER(mm128_v), C.c_char_p, i32]
    L.mm_sketch2.restype = mm_pattern_t
    L.mm_sketch2.argtypes = [vp, C.c_char_p, i32, i32, i32, C.c_uint32, i32, C.POINTER(mm128_v), C.c_char_p, i32, C.c_float]
    L.mm_sketch3.restype = C.c_uint
    L.mm_sketch3.argtypes = [vp, C.c_char_p, C.c_uint, i32, i32, C.c_uint32, i32, C.POINTER(mm128_v), C.c_char_p, i32, i32,
                             C.c_uint32]
    L.gd_sketch_ref_batch.restype = i32
    L.gd_sketch_ref_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, C.c_char_p, i32, vp, vp, i64]
    L.gd_sketch_ref_batch_device.restype = i32
    L.gd_sketch_ref_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, i64, i32, i32, C.c_char_p, i32, vp, vp, i64]
    L.gd_sketch_reads_batch.restype = i32
    L.gd_sketch_reads_batch.argtypes = [vp, i32, vp, vp, vp, i32, i32, C.c_char_p, i32, C.c_float, C.c_uint32, vp, vp, vp,
                                        i64, vp, vp, vp, i64]
    L.gd_index_build.restype = i32
    L.gd_index_build.argtypes = [vp, i32, vp, vp, vp, i32, i32, C.c_char_p, i32, C.POINTER(vp)]
    L.gd_index_build_device.restype = i32
    L.gd_index_build_device.argtypes = [vp, i32, vp, vp, vp, i32, i32, C.c_char_p, i32, C.POINTER(vp)]
    L.gd_index_destroy.restype = None
    L.gd_index_destroy.argtypes = [vp]
    L.gd_index_stat.restype = i64
    L.gd_index_stat.argtypes = [vp, C.c_char_p]
    L.gd_index_get_batch.restype = i32
    L.gd_index_get_batch.argtypes = [vp, vp, i64, vp, vp, vp]
    L.gd_index_export.restype = i32
    L.gd_index_export.argtypes = [vp, vp, vp, vp, vp, vp]
    L.gd_index_cal_max_occ.restype = i32
    L.gd_index_cal_max_occ.argtypes = [vp, vp, C.c_float, C.POINTER(C.c_int32)]
    L.gd_sr_map_batch.restype = i32
    L.gd_sr_map_batch.argtypes = [vp, vp, i32, vp, vp, vp, C.POINTER(gd_sr_opt_t), vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.gd_index_meta.restype = i32
    L.gd_index_meta.argtypes = [vp, C.POINTER(gd_index_meta_t)]
    L.gd_index_alloc.restype = i32
    L.gd_index_alloc.argtypes = [vp, C.POINTER(gd_index_meta_t), C.POINTER(vp)]
    L.gd_index_buffers.restype = i32
    L.gd_index_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.gd_index_commit.restype = i32
    L.gd_index_commit.argtypes = [vp, vp]
    L.gd_lr_map_batch.restype = i32
    L.gd_lr_map_batch.argtypes = [vp, vp, i32, vp, vp, vp, C.POINTER(gd_lr_opt_t), vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.gd_index_load_mmi.restype = i32
    L.gd_index_load_mmi.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.gd_index_seq_name.restype = C.c_char_p
    L.gd_index_seq_name.argtypes = [vp, i32]
    L.gd_mmi_write.restype = i32
    L.gd_mmi_write.argtypes = [C.c_char_p, i32, i32, i32, i32, i32, vp, vp, i64, vp, vp, vp, vp]
    L.gd_sr_sam_batch.restype = i32
    L.gd_sr_sam_batch.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(gd_sr_post_opt_t),
                                  C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.gd_pinned_alloc.restype = vp
    L.gd_pinned_alloc.argtypes = [C.c_size_t]
    L.gd_pinned_free.restype = None
    L.gd_pinned_free.argtypes = [vp]
    L.gd_multi_init.restype = i32
    L.gd_multi_init.argtypes = [i32, vp, C.POINTER(vp)]
    L.gd_multi_destroy.restype = None
    L.gd_multi_destroy.argtypes = [vp]
    L.gd_multi_size.restype = i32
    L.gd_multi_size.argtypes = [vp]
    L.gd_multi_ctx.restype = vp
    L.gd_multi_ctx.argtypes = [vp, i32]
    L.gd_multi_index.restype = vp
    L.gd_multi_index.argtypes = [vp, i32]
    L.gd_multi_strerror.restype = C.c_char_p
    L.gd_multi_strerror.argtypes = [vp]
    L.gd_multi_index_bcast.restype = i32
    L.gd_multi_index_bcast.argtypes = [vp, vp, i32]
    L.gd_multi_stat.restype = C.c_double
    L.gd_multi_stat.argtypes = [vp, C.c_char_p]
    L.gd_multi_sr_map_batch.restype = i32
    L.gd_multi_sr_map_batch.argtypes = [vp, i32, vp, vp, vp, C.POINTER(gd_sr_opt_t), vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.gd_multi_lr_map_batch.restype = i32
    L.gd_multi_lr_map_batch.argtypes = [vp, i32, vp, vp, vp, C.POINTER(gd_lr_opt_t), vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.gd_sr_map_sam_batch.restype = i32
    L.gd_sr_map_sam_batch.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, C.POINTER(gd_sr_opt_t), C.POINTER(gd_sr_post_opt_t), i32, vp,
                                      C.POINTER(C.POINTER(vp)), C.POINTER(C.POINTER(C.c_size_t)), C.POINTER(i32)]
    for fn, ot in ((L.gd_multi_sr_map_sam, gd_sr_opt_t), (L.gd_multi_lr_map_sam, gd_lr_opt_t)):
        fn.restype = i32
        fn.argtypes = [vp, i32, vp, vp, vp, vp, vp, C.POINTER(ot), C.POINTER(gd_sr_post_opt_t), i32, vp, vp, vp, vp,
                       C.POINTER(C.POINTER(vp)), C.POINTER(C.POINTER(C.c_size_t)), C.POINTER(i32)]
    L.gd_sr_sam_batch_parts.restype = i32
    L.gd_sr_sam_batch_parts.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(gd_sr_post_opt_t),
                                        C.POINTER(C.POINTER(vp)), C.POINTER(C.POINTER(C.c_size_t)), C.POINTER(i32)]
    L.gd_lr_sam_batch.restype = i32
    L.gd_lr_sam_batch.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(gd_sr_post_opt_t),
                                  C.POINTER(vp), C.POINTER(C.c_size_t), vp, vp]
    L.gd_sam_header.restype = i32
    L.gd_sam_header.argtypes = [i32, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.gd_free.restype = None
    L.gd_free.argtypes = [vp]
    _lib = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):  # torch tensor (device or host)
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def score_matrix(a, b):
    """5x5 matrix as the live call site builds it (GDiet-ShortReads/map.c:861-865)."""
    bb = -abs(int(b))
    m = np.full((5, 5), bb, np.int8)
    np.fill_diagonal(m, a)
    m[4, :] = 0
    m[:, 4] = 0
    return np.ascontiguousarray(m.reshape(-1))


class KswParams:
    """Batch-uniform arguments of ksw_extd2_sse (GDiet-ShortReads/ksw2.h:42-59)."""

    def __init__(self, mat, q, e, q2, e2, zdrop, end_bonus, flag, m=None):
        self.mat = np.ascontiguousarray(mat, np.int8)
        self.m = int(m if m is not None else round(len(self.mat) ** 0.5))
        self.c = gd_ksw_params_t(self.m, self.mat.ctypes.data, int(q), int(e), int(q2), int(e2), int(zdrop),
                                 int(end_bonus), int(flag))
        self.flag = int(flag)


class Context:
    """gd_ctx: one CUDA stream + staging memory. Raises GdietError when no sm_100 GPU is usable."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.gd_init(device, C.byref(h))
        if rc != GD_OK:
            raise GdietError("gd_init failed (%d): %s" % (rc, self.lib.gd_strerror(None).decode()))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_owned", True):  # (views handed out by Multi.ctx() belong to the gd_multi)
                self.lib.gd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != GD_OK:
            raise GdietError("%s failed (%d): %s" % (what, rc, self.lib.gd_strerror(self.h).decode()))

    def set_option(self, key, value):
        self._check(self.lib.gd_set_option(self.h, key.encode(), int(value)), "gd_set_option")

    def stat(self, key):
        return int(self.lib.gd_get_stat(self.h, key.encode()))

    @property
    def stream(self):
        return int(self.lib.gd_stream(self.h) or 0)

    # ---- DP ----
    def ksw_extd2_batch(self, qlen, qoff, qbuf, tlen, toff, tbuf, params, w=None, w_all=-1, want_cigar=True,
                        out=None):
        """gd_ksw_extd2_batch over host arrays (numpy, or pinned torch tensors). Returns (ez, cigar_off, cigar)."""
        n = len(qlen)
        ez = np.zeros(n, GD_EXTZ_DTYPE) if out is None else out["ez"]
        coff = np.zeros(n + 1, np.int64) if out is None else out["cigar_off"]
        if want_cigar:
            if out is None:
                cap = int(n) * 64 + 1024
                cig = np.zeros(cap, np.uint32)
            else:
                cig = out["cigar"]
                cap = len(cig)
        else:
            cig, cap = None, 0
        while True:
            rc = self.lib.gd_ksw_extd2_batch(self.h, n, _ptr(qlen), _ptr(qoff), _ptr(qbuf), _ptr(tlen), _ptr(toff),
                                             _ptr(tbuf), _ptr(w), int(w_all), C.byref(params.c), _ptr(ez),
                                             _ptr(coff), _ptr(cig), cap)
            if rc == GD_ERR_CAPACITY and want_cigar and out is None and int(coff[n]) > cap:
                cap = int(coff[n]) + 16
                cig = np.zeros(cap, np.uint32)
                continue
            self._check(rc, "gd_ksw_extd2_batch")
            break
        return ez, coff, (cig[: int(coff[n])] if cig is not None and out is None else cig)

    def ksw_extd2_batch_device(self, n, d_qlen, d_qoff, d_qbuf, d_tlen, d_toff, d_tbuf, params, max_qlen, max_tlen,
                               max_w, d_ez, d_cigar=None, cigar_stride=0, d_w=None, w_all=-1):
        """gd_ksw_extd2_batch_device: all arrays are device pointers / torch CUDA tensors; enqueues only."""
        rc = self.lib.gd_ksw_extd2_batch_device(self.h, n, _ptr(d_qlen), _ptr(d_qoff), _ptr(d_qbuf), _ptr(d_tlen),
                                                _ptr(d_toff), _ptr(d_tbuf), _ptr(d_w), int(w_all), int(max_qlen),
                                                int(max_tlen), int(max_w), C.byref(params.c), _ptr(d_ez),
                                                _ptr(d_cigar), int(cigar_stride))
        self._check(rc, "gd_ksw_extd2_batch_device")

    def exact_match_batch_device(self, n, d_qlen, d_qoff, d_qbuf, d_toff, d_tbuf, d_equal):
        self._check(self.lib.gd_exact_match_batch_device(self.h, n, _ptr(d_qlen), _ptr(d_qoff), _ptr(d_qbuf),
                                                         _ptr(d_toff), _ptr(d_tbuf), _ptr(d_equal)),
                    "gd_exact_match_batch_device")

    # ---- sketching ----
    def sketch_ref_batch(self, off, lens, buf, w, k, Z, rid=None, out_cap=None):
        """gd_sketch_ref_batch: mm_sketch of n sequences. Returns (out_off[n+1], mm128 array)."""
        n = len(lens)
        Zb = Z.encode() if isinstance(Z, str) else Z
        out_off = np.zeros(n + 1, np.int64)
        if out_cap is None:
            out_cap = int(np.sum(lens)) // max(1, len(Zb)) * Zb.count(b"1") // max(1, w // 2) + 4096
        while True:
            out = np.zeros(out_cap, MM128_DTYPE)
            rc = self.lib.gd_sketch_ref_batch(self.h, n, _ptr(off), _ptr(lens), _ptr(rid), _ptr(buf), w, k, Zb, len(Zb),
                                              _ptr(out_off), _ptr(out), out_cap)
            if rc == GD_ERR_CAPACITY and int(out_off[n]) > out_cap:
                out_cap = int(out_off[n]) + 16
                continue
            self._check(rc, "gd_sketch_ref_batch")
            return out_off, out[: int(out_off[n])]

    def sketch_ref_batch_device(self, n, d_off, d_len, d_rid, d_buf, total_len, w, k, Z, d_out_off, d_out, out_cap):
        Zb = Z.encode() if isinstance(Z, str) else Z
        self._check(self.lib.gd_sketch_ref_batch_device(self.h, n, _ptr(d_off), _ptr(d_len), _ptr(d_rid), _ptr(d_buf),
                                                        int(total_len), w, k, Zb, len(Zb), _ptr(d_out_off), _ptr(d_out),
                                                        int(out_cap)), "gd_sketch_ref_batch_device")

    def sketch_reads_batch(self, off, lens, buf, w, k, Z, max_seeds, max_nb_seeds):
        """gd_sketch_reads_batch. Returns dict(s2_counts[n,W], s2_off[n+1], s2, s3_off[n*W+1], s3_ret[n,W], s3)."""
        n = len(lens)
        Zb = Z.encode() if isinstance(Z, str) else Z
        W = len(Zb)
        s2_counts = np.zeros((n, W), np.uint32)
        s3_ret = np.zeros((n, W), np.uint32)
        s2_off = np.zeros(n + 1, np.int64)
        s3_off = np.zeros(n * W + 1, np.int64)
        cap = int(np.sum(lens)) // max(1, w // 2) * W + 4096 * W
        while True:
            s2 = np.zeros(cap, MM128_DTYPE)
            s3 = np.zeros(cap, MM128_DTYPE)
            rc = self.lib.gd_sketch_reads_batch(self.h, n, _ptr(off), _ptr(lens), _ptr(buf), w, k, Zb, W,
                                                float(max_seeds), int(max_nb_seeds), _ptr(s2_counts), _ptr(s2_off),
                                                _ptr(s2), cap, _ptr(s3_off), _ptr(s3_ret), _ptr(s3), cap)
            if rc == GD_ERR_CAPACITY:
                cap *= 4
                continue
            self._check(rc, "gd_sketch_reads_batch")
            return dict(s2_counts=s2_counts, s2_off=s2_off, s2=s2[: int(s2_off[n])], s3_off=s3_off, s3_ret=s3_ret,
                        s3=s3[: int(s3_off[n * W])])


    # ---- index + short-read mapping ----
    def index_build(self, contigs, w, k, Z):
        """gd_index_build over a list of ASCII contigs (uint8 arrays / bytes). Returns an Index."""
        bufs = [np.frombuffer(c, np.uint8) if isinstance(c, (bytes, bytearray)) else np.ascontiguousarray(c, np.uint8) for c in contigs]
        lens = np.array([len(b) for b in bufs], np.int32)
        off = np.zeros(len(bufs), np.int64)
        off[1:] = np.cumsum(lens[:-1].astype(np.int64))
        buf = np.concatenate(bufs) if len(bufs) > 1 else bufs[0]
        Zb = Z.encode() if isinstance(Z, str) else Z
        h = C.c_void_p()
        self._check(self.lib.gd_index_build(self.h, len(bufs), _ptr(off), _ptr(lens), _ptr(buf), w, k, Zb, len(Zb), C.byref(h)),
                    "gd_index_build")
        return Index(self, h)

    def index_load_mmi(self, path):
        """gd_index_load_mmi: a .mmi file (the reference's `-d` output or gd_mmi_write's) -> device index."""
        h = C.c_void_p()
        self._check(self.lib.gd_index_load_mmi(self.h, path.encode(), C.byref(h)), "gd_index_load_mmi")
        return Index(self, h)

    def index_build_device(self, off, lens, d_buf, w, k, Z):
        """gd_index_build_device: off / lens are host arrays, d_buf a device pointer / CUDA tensor with the ASCII contigs."""
        off = np.ascontiguousarray(off, np.int64)
        lens = np.ascontiguousarray(lens, np.int32)
        Zb = Z.encode() if isinstance(Z, str) else Z
        h = C.c_void_p()
        self._check(self.lib.gd_index_build_device(self.h, len(lens), _ptr(off), _ptr(lens), _ptr(d_buf), w, k, Zb, len(Zb), C.byref(h)),
                    "gd_index_build_device")
        return Index(self, h)

    def lr_map_batch(self, index, off, lens, buf, opt, cand_cap=None, cigar_cap=None):
        """gd_lr_map_batch (long-read tree). Returns (cand_off[n+1], candidates (SR_CAND_DTYPE), cigar pool)."""
        return self.sr_map_batch(index, off, lens, buf, opt, cand_cap, cigar_cap, fn=self.lib.gd_lr_map_batch, what="gd_lr_map_batch")

    def sr_map_sam_batch(self, index, names, off, lens, seq, qual, opt, post, seq_names, join=True):
        """gd_sr_map_sam_batch: reads in, SAM text out (the post-DP stage runs on the device).  names / seq_names: lists of
        str or ctypes char* arrays.  Returns the text (bytes), or with join=False the list of (address, length) pieces, which
        stay valid until the call after the next one."""
        n_arr, s_arr = _cstr_array(names), _cstr_array(seq_names)
        parts, plen, npart = C.POINTER(C.c_void_p)(), C.POINTER(C.c_size_t)(), C.c_int(0)
        self._check(self.lib.gd_sr_map_sam_batch(self.h, index.h, len(lens), C.cast(n_arr, C.c_void_p), _ptr(off), _ptr(lens), _ptr(seq),
                                                 _ptr(qual), C.byref(opt), C.byref(post), len(s_arr), C.cast(s_arr, C.c_void_p),
                                                 C.byref(parts), C.byref(plen), C.byref(npart)), "gd_sr_map_sam_batch")
        pieces = [(int(parts[i] or 0), int(plen[i])) for i in range(npart.value)]
        self.lib.gd_free(C.cast(parts, C.c_void_p)), self.lib.gd_free(C.cast(plen, C.c_void_p))
        if not join:
            return pieces
        return b"".join(C.string_at(a, l) for a, l in pieces)

    def sr_map_batch(self, index, off, lens, buf, opt, cand_cap=None, cigar_cap=None, fn=None, what="gd_sr_map_batch"):
        """gd_sr_map_batch. Returns (cand_off[n+1], candidates (SR_CAND_DTYPE), cigar pool)."""
        fn = fn or self.lib.gd_sr_map_batch
        n = len(lens)
        cand_off = np.zeros(n + 1, np.int64)
        cand_cap = cand_cap or 2 * n + 64
        cigar_cap = cigar_cap or 16 * n + 1024
        ncig = C.c_int64(0)
        while True:
            cand = np.zeros(cand_cap, SR_CAND_DTYPE)
            cig = np.zeros(cigar_cap, np.uint32)
            rc = fn(self.h, index.h, n, _ptr(off), _ptr(lens), _ptr(buf), C.byref(opt), _ptr(cand_off), _ptr(cand), cand_cap,
                    _ptr(cig), cigar_cap, C.byref(ncig))
            if rc == GD_ERR_CAPACITY:
                cand_cap = max(cand_cap, int(cand_off[n]) + 16)
                cigar_cap = max(cigar_cap, int(ncig.value) + 16)
                continue
            self._check(rc, what)
            return cand_off, cand[: int(cand_off[n])], cig[: int(ncig.value)]


class Multi:
    """gd_multi: several GPUs of one box from one process -- index broadcast, contiguous read shards, input-order results."""

    def __init__(self, n_dev):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.gd_multi_init(int(n_dev), None, C.byref(h))
        if rc != GD_OK:
            raise GdietError("gd_multi_init(%d) failed: %s" % (n_dev, self.lib.gd_strerror(None).decode()))
        self.h, self.n = h, int(n_dev)

    def _check(self, rc, what):
        if rc != GD_OK:
            raise GdietError("%s failed (%d): %s" % (what, rc, self.lib.gd_multi_strerror(self.h).decode()))

    def ctx(self, i):
        """a non-owning Context view of device i's gd_ctx"""
        c = Context.__new__(Context)
        c.lib, c.h, c.device, c._owned = self.lib, C.c_void_p(self.lib.gd_multi_ctx(self.h, i)), i, False
        return c

    def index_bcast(self, index):
        """index lives on ctx(0); the handle belongs to the Multi afterwards"""
        self._check(self.lib.gd_multi_index_bcast(self.h, index.h, 1), "gd_multi_index_bcast")
        index.h = None
        return {k: self.lib.gd_multi_stat(self.h, k.encode()) for k in ("bcast_seconds", "bcast_bytes", "bcast_path")}

    def index(self, i):
        return Index(self.ctx(i), C.c_void_p(self.lib.gd_multi_index(self.h, i)), owned=False)

    def map_batch(self, off, lens, buf, opt, cand_cap=None, cigar_cap=None):
        """gd_multi_sr_map_batch / gd_multi_lr_map_batch (by the type of opt): same results as one device."""
        fn = self.lib.gd_multi_lr_map_batch if isinstance(opt, gd_lr_opt_t) else self.lib.gd_multi_sr_map_batch
        n = len(lens)
        cand_off = np.zeros(n + 1, np.int64)
        cand_cap = cand_cap or 2 * n + 64
        cigar_cap = cigar_cap or 16 * n + 1024
        ncig = C.c_int64(0)
        while True:
            cand = np.zeros(cand_cap, SR_CAND_DTYPE)
            cig = np.zeros(cigar_cap, np.uint32)
            rc = fn(self.h, n, _ptr(off), _ptr(lens), _ptr(buf), C.byref(opt), _ptr(cand_off), _ptr(cand), cand_cap, _ptr(cig), cigar_cap,
                    C.byref(ncig))
            if rc == GD_ERR_CAPACITY:
                cand_cap, cigar_cap = max(cand_cap, int(cand_off[n]) + 16), max(cigar_cap, int(ncig.value) + 16)
                continue
            self._check(rc, "gd_multi_map_batch")
            return cand_off, cand[: int(cand_off[n])], cig[: int(ncig.value)]

    def map_sam(self, names, off, lens, seq, qual, opt, post, seq_names, contigs):
        """gd_multi_sr_map_sam / gd_multi_lr_map_sam: the SAM text of the batch (pieces joined in input order)."""
        fn = self.lib.gd_multi_lr_map_sam if isinstance(opt, gd_lr_opt_t) else self.lib.gd_multi_sr_map_sam
        bufs = [np.ascontiguousarray(c, np.uint8) for c in contigs]
        rlen = np.array([len(b) for b in bufs], np.int32)
        roff = np.zeros(len(bufs), np.int64)
        roff[1:] = np.cumsum(rlen[:-1].astype(np.int64))
        ref = np.concatenate(bufs) if len(bufs) > 1 else bufs[0]
        parts, plen, npart = C.POINTER(C.c_void_p)(), C.POINTER(C.c_size_t)(), C.c_int(0)
        n_arr, s_arr = _cstr_array(names), _cstr_array(seq_names)
        self._check(fn(self.h, len(lens), C.cast(n_arr, C.c_void_p), _ptr(off), _ptr(lens), _ptr(seq), _ptr(qual), C.byref(opt), C.byref(post),
                       len(bufs), C.cast(s_arr, C.c_void_p), _ptr(roff), _ptr(rlen), _ptr(ref), C.byref(parts), C.byref(plen), C.byref(npart)),
                    "gd_multi_map_sam")
        out = [C.string_at(parts[i], plen[i]) for i in range(npart.value)]  # (the pieces belong to the handle)
        self.lib.gd_free(C.cast(parts, C.c_void_p)), self.lib.gd_free(C.cast(plen, C.c_void_p))
        return b"".join(out)

    def close(self):
        if self.h:
            self.lib.gd_multi_destroy(self.h)
            self.h = None


class Index:
    """gd_index: the device-resident minimizer index (mm_idx_t of GDiet-ShortReads/minimap.h:86-96)."""

    def __init__(self, ctx, h, owned=True):
        self.ctx, self.h, self.owned = ctx, h, owned

    def stat(self, key):
        return int(self.ctx.lib.gd_index_stat(self.h, key.encode()))

    def get(self, minier):
        """mm_idx_get for an array of minimizer values (x >> 8): (count, first position offset or -1)."""
        m = np.ascontiguousarray(minier, np.uint64)
        cnt = np.zeros(len(m), np.uint32)
        first = np.zeros(len(m), np.int64)
        self.ctx._check(self.ctx.lib.gd_index_get_batch(self.ctx.h, self.h, len(m), _ptr(m), _ptr(cnt), _ptr(first)),
                        "gd_index_get_batch")
        return cnt, first

    def export(self):
        keys = np.zeros(self.stat("n_keys"), np.uint64)
        counts = np.zeros(self.stat("n_keys"), np.uint32)
        pos = np.zeros(self.stat("n_minimizers"), np.uint64)
        S = np.zeros(self.stat("s_words"), np.uint32)
        self.ctx._check(self.ctx.lib.gd_index_export(self.ctx.h, self.h, _ptr(keys), _ptr(counts), _ptr(pos), _ptr(S)),
                        "gd_index_export")
        return keys, counts, pos, S

    def seq_names(self):
        return [self.ctx.lib.gd_index_seq_name(self.h, i).decode() for i in range(self.stat("n_seq"))]

    def meta(self):
        m = gd_index_meta_t()
        self.ctx._check(self.ctx.lib.gd_index_meta(self.h, C.byref(m)), "gd_index_meta")
        return m

    def buffers(self):
        """[(device pointer, bytes)] of the GD_INDEX_NBUF device buffers."""
        ptrs, nb = (C.c_void_p * GD_INDEX_NBUF)(), (C.c_size_t * GD_INDEX_NBUF)()
        self.ctx._check(self.ctx.lib.gd_index_buffers(self.h, ptrs, nb), "gd_index_buffers")
        return [(int(ptrs[i] or 0), int(nb[i])) for i in range(GD_INDEX_NBUF)]

    @staticmethod
    def alloc(ctx, meta):
        h = C.c_void_p()
        ctx._check(ctx.lib.gd_index_alloc(ctx.h, C.byref(meta), C.byref(h)), "gd_index_alloc")
        return Index(ctx, h)

    def commit(self):
        self.ctx._check(self.ctx.lib.gd_index_commit(self.ctx.h, self.h), "gd_index_commit")

    def cal_max_occ(self, frac):
        v = C.c_int32(0)
        self.ctx._check(self.ctx.lib.gd_index_cal_max_occ(self.ctx.h, self.h, float(frac), C.byref(v)), "gd_index_cal_max_occ")
        return int(v.value)

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "owned", True):
                self.ctx.lib.gd_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# drop-in single-call mirrors: these call the symbols that carry the REFERENCE's names
# ---------------------------------------------------------------------------------------------
def ksw_extd2(query, target, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, entry="ksw_extd2_avx512", m=None):
    """Calls the drop-in ``ksw_extd2_avx512`` / ``ksw_extd2_sse`` symbol of libgdiet_cuda.so exactly like
    GDiet-ShortReads/map.c:923-929 does. Returns (dict of ksw_extz_t fields, cigar array)."""
    L = load()
    query = np.ascontiguousarray(query, np.uint8)
    target = np.ascontiguousarray(target, np.uint8)
    mat = np.ascontiguousarray(mat, np.int8)
    if m is None:
        m = int(round(len(mat) ** 0.5))
    ez = ksw_extz_t()  # memset 0 like map.c:866
    getattr(L, entry)(None, len(query), _ptr(query), len(target), _ptr(target), m, _ptr(mat), q, e, q2, e2, w, zdrop,
                      end_bonus, flag, C.byref(ez))
    d = dict(max=ez.max_zdropped & 0x7fffffff, zdropped=ez.max_zdropped >> 31, max_q=ez.max_q, max_t=ez.max_t,
             mqe=ez.mqe, mqe_t=ez.mqe_t, mte=ez.mte, mte_q=ez.mte_q, score=ez.score, n_cigar=ez.n_cigar,
             reach_end=ez.reach_end)
    cig = np.array([ez.cigar[i] for i in range(ez.n_cigar)], np.uint32)
    if ez.cigar:
        _libc.free(C.cast(ez.cigar, C.c_void_p))  # kfree(km=NULL, ...) == free (kalloc.c)
    return d, cig


def _take_v(v):
    out = np.zeros(v.n, MM128_DTYPE)
    if v.n:
        C.memmove(out.ctypes.data, v.a, v.n * 16)
    if v.a:
        _libc.free(v.a)
    return np.stack([out["x"], out["y"]], 1) if v.n else np.zeros((0, 2), np.uint64)


def mm_sketch(seq, w, k, rid, Z):
    """drop-in mm_sketch (GDiet-ShortReads/mmpriv.h:63)"""
    L = load()
    Zb = Z.encode() if isinstance(Z, str) else Z
    v = mm128_v(0, 0, None)
    L.mm_sketch(None, bytes(seq), len(seq), w, k, rid, 0, C.byref(v), Zb, len(Zb))
    return _take_v(v)


def mm_sketch3(seq, w, k, rid, Z, shift, max_nb_seeds):
    """drop-in mm_sketch3 (GDiet-ShortReads/mmpriv.h:67). Returns (entries, return value)."""
    L = load()
    Zb = Z.encode() if isinstance(Z, str) else Z
    v = mm128_v(0, 0, None)
    ret = L.mm_sketch3(None, bytes(seq), len(seq), w, k, rid, 0, C.byref(v), Zb, len(Zb), shift, max_nb_seeds)
    return _take_v(v), int(ret)


def mm_sketch2(seq, w, k, rid, Z, max_seeds):
    """drop-in mm_sketch2 (GDiet-ShortReads/mmpriv.h:65). Returns (entries, shift_seeds_number)."""
    L = load()
    Zb = Z.encode() if isinstance(Z, str) else Z
    v = mm128_v(0, 0, None)
    pat = L.mm_sketch2(None, bytes(seq), len(seq), w, k, rid, 0, C.byref(v), Zb, len(Zb), float(max_seeds))
    counts = np.array([pat.shift_seeds_number[i] for i in range(pat.n)], np.uint32)
    _libc.free(C.cast(pat.shift_seeds_number, C.c_void_p))
    return _take_v(v), counts
