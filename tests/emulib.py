"""Loader for tests/emu/libgd_emu.so: the DEVICE CODE of csrc/gd_ksw.cuh / gd_sketch.cuh compiled for
the host against the fiber SIMT emulator (tests/emu/simt_emu.h). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from oraclelib import EXTZ_FIELDS, u8p, i8p, i32p, i64p, u32p, u64p

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(ROOT, "genome-on-diet_b200", "csrc")
RES = np.dtype([(f, np.int32) for f in EXTZ_FIELDS + ["tb_i", "tb_j", "rows_done", "lead64", "pad1"]])


def build():
    so = os.path.join(EMU_DIR, "libgd_emu.so")
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_ksw.cpp", "emu_sketch.cpp", "emu_sam.cpp")]
    deps = srcs + [os.path.join(EMU_DIR, "simt_emu.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                              if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", EMU_DIR, "-I", CSRC, "-o", so] + srcs)
    return so


class Emu:
    def __init__(self):
        L = self.lib = C.CDLL(build())
        L.emu_ksw_batch.restype = C.c_int
        L.emu_ksw_batch.argtypes = [C.c_int, i32p, i64p, u8p, i32p, i64p, u8p, i32p, C.c_int, i8p] + [C.c_int] * 9 + [
            C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.emu_sam_batch.restype = C.c_int
        L.emu_sam_batch.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.c_int64, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_void_p),
                                                                                                         C.POINTER(C.c_size_t)]
        L.emu_sam_check_fixed4.restype = C.c_long
        L.emu_sam_check_fixed4.argtypes = [C.c_int]
        L.emu_sketch_jobs.restype = C.c_long
        L.emu_sketch_jobs.argtypes = [C.c_int, i64p, i32p, i32p, u32p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                      C.c_int, C.c_int, i64p, u64p, C.c_int64]

    def sketch_concurrency(self, seed):
        """seed != 0: the sketch tiles run with all blocks resident, interleaved pseudo-randomly (simt_emu.h: launch_concurrent)"""
        self.lib.emu_sketch_concurrency.argtypes = [C.c_uint]
        self.lib.emu_sketch_concurrency(seed)

    def sketch_packed(self, seqs, shifts, rids, w, k, Z, pack, threads, grid=3):
        """fixed-stride output, `pack` whole jobs per tile (0: one job per one-warp tile); returns the list of every job"""
        L = self.lib
        L.emu_sketch_packed.restype = C.c_int
        L.emu_sketch_packed.argtypes = [C.c_int, i64p, i32p, i32p, u32p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int64, i32p, u64p]
        buf = b"".join(seqs)
        lens = np.array([len(s) for s in seqs], np.int32)
        off = np.zeros(len(seqs), np.int64)
        off[1:] = np.cumsum(lens[:-1])
        stride = int(lens.max()) + 2
        out = np.zeros(2 * stride * len(seqs), np.uint64)
        cnt = np.full(len(seqs), -1, np.int32)
        rc = L.emu_sketch_packed(len(seqs), off, lens, np.array(shifts, np.int32), np.array(rids, np.uint32), buf, w, k, Z.encode(),
                                 len(Z), pack, threads, grid, stride, cnt, out)
        assert rc == 0, rc
        return [out[2 * stride * i:2 * (stride * i + cnt[i])].reshape(-1, 2) for i in range(len(seqs))]

    def ksw_batch(self, P, w, mat, sc, flag, G, threads=64):
        n = P["n"]
        res = np.zeros(n, RES)
        stride = int((P["qlen"] + P["tlen"]).max()) + 8
        cig = np.zeros(n * stride, np.uint32)
        l64 = C.c_int(0)
        rc = self.lib.emu_ksw_batch(n, P["qlen"], P["qoff"], P["qbuf"], P["tlen"], P["toff"], P["tbuf"],
                                    np.ascontiguousarray(w, np.int32), 5, mat, sc["q"], sc["e"], sc["q2"], sc["e2"],
                                    sc["zdrop"], sc["end_bonus"], flag, G, threads, res.ctypes.data_as(C.c_void_p),
                                    cig.ctypes.data_as(C.c_void_p), stride, C.byref(l64))
        assert rc == 0
        self.last_lead64 = int(l64.value)
        return res, cig.reshape(n, stride)

    def sam_batch(self, names, off, lens, seq, qual, cand_off, cand, cigar, seq_names, contigs, opt):
        """csrc/gd_sam_core.h (the per-read SAM stage of the GPU path) run on the host; same arguments as gd.sr_sam_batch"""
        import gdiet_b200 as gd
        ref_len = np.array([len(c) for c in contigs], np.int32)
        ref_off = np.zeros(len(contigs), np.int64)
        ref_off[1:] = np.cumsum(ref_len[:-1].astype(np.int64))
        ref = np.concatenate([np.ascontiguousarray(c, np.uint8) for c in contigs])
        n_arr, s_arr = gd._cstr_array(names), gd._cstr_array(seq_names)
        cand = np.ascontiguousarray(cand) if len(cand) else np.zeros(1, gd.SR_CAND_DTYPE)
        cig = np.ascontiguousarray(cigar, np.uint32) if len(cigar) else np.zeros(1, np.uint32)
        out, out_len = C.c_void_p(), C.c_size_t(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.emu_sam_batch(len(lens), C.cast(n_arr, C.c_void_p), p(off), p(lens), p(seq), p(qual), p(cand_off), p(cand), p(cig),
                                    len(cigar), len(contigs), C.cast(s_arr, C.c_void_p), p(ref_off), p(ref_len), p(ref), C.byref(opt),
                                    C.byref(out), C.byref(out_len))
        assert rc == 0, rc
        txt = C.string_at(out, out_len.value)
        self.lib.emu_sam_check_fixed4.argtypes  # (keep the handle alive)
        C.CDLL(None).free(out)
        return txt

    def sketch_jobs(self, seqs, shifts, rids, w, k, Z, small, grid=3):
        buf = b"".join(seqs)
        lens = np.array([len(s) for s in seqs], np.int32)
        off = np.zeros(len(seqs), np.int64)
        off[1:] = np.cumsum(lens[:-1])
        cap = len(buf) + 16
        out = np.zeros(2 * cap, np.uint64)
        oo = np.zeros(len(seqs) + 1, np.int64)
        self.lib.emu_sketch_jobs(len(seqs), off, lens, np.array(shifts, np.int32), np.array(rids, np.uint32), buf, w, k,
                                 Z.encode(), len(Z), small, grid, oo, out, cap)
        return [out[2 * oo[i]:2 * oo[i + 1]].reshape(-1, 2) for i in range(len(seqs))]
