"""CPU: the device code of the CUDA kernels, compiled for the host against a fiber SIMT emulator, must
reproduce the oracle.  This checks lane ownership, the shared-memory ring, carries between steps, the
tag-carrying packed arithmetic, the exact-max reduction and the look-back compaction without a GPU.
(The GPU parity tests in test_gpu_*.py remain the authority for the real hardware path.)"""
import ctypes as C
import os

import numpy as np
import pytest

import gdiet_b200  # noqa: F401
from gdiet_b200 import synth
from oraclelib import EXTZ_FIELDS
from helpers import pair


@pytest.fixture(scope="module")
def emu():
    from emulib import Emu
    return Emu()


def _check(emu, oracle, P, w, scn, flag, G, force_literal_max=False):
    sc = synth.SCORING[scn] if isinstance(scn, str) else scn
    mat = synth.score_matrix(sc["a"], sc["b"])
    res, cig = emu.ksw_batch(P, w, mat, sc, flag | (0x100 if force_literal_max else 0), G)
    for i in range(P["n"]):
        q, t = pair(P, i)
        ez, oc = oracle.ksw_extd2(q, t, mat, sc["q"], sc["e"], sc["q2"], sc["e2"], int(w[i]), sc["zdrop"], sc["end_bonus"], flag)
        mine = {f: int(res[i][f]) for f in EXTZ_FIELDS}
        assert mine == ez, "pair %d flag %#x G %d" % (i, flag, G)
        assert np.array_equal(cig[i][:ez["n_cigar"]], oc), "pair %d flag %#x G %d" % (i, flag, G)


@pytest.mark.parametrize("flag", [0x08, 0x00, 0x18, 0x0a, 0x40, 0x42, 0x09, 0x80])
def test_emu_ksw_ragged(emu, oracle, flag):
    P = synth.ragged_pairs(10, seed=5 + flag, max_len=200)
    rng = np.random.default_rng(flag)
    for G, scn in ((4, "sr"), (8, "map-hifi"), (16, "map-ont"), (32, "sr")):
        w = rng.choice([-1, 5, 10, 20, 33, 37, 100, 150, 400], P["n"]).astype(np.int32)
        _check(emu, oracle, P, w, scn, flag, G)


@pytest.mark.parametrize("flag", [0x18, 0x58])
def test_emu_ksw_approx_drop_on_last_row(emu, oracle, flag):
    """KSW_EZ_APPROX_DROP: a Z-drop on the last anti-diagonal leaves the loop before ez->score is set
    (ksw2_extd2_sse.c:380-382); tiny pairs and small Z-drop values reach that row (found by tools/ksw_fuzz.py)"""
    P = synth.ragged_pairs(60, seed=9, max_len=8)
    w = np.random.default_rng(1).choice([-1, 1, 5, 33], P["n"]).astype(np.int32)
    for zdrop in (0, 5):
        sc = dict(synth.SCORING["sr"], zdrop=zdrop)
        for G in (4, 32):
            _check(emu, oracle, P, w, sc, flag, G)


@pytest.mark.parametrize("G,wv", [(4, 10), (8, 37), (32, 64), (32, 5)])
def test_emu_ksw_ring_wrap(emu, oracle, G, wv):
    """narrow uniform band on longer pairs: the column ring wraps many times"""
    P = synth.ragged_pairs(8, seed=9, max_len=700)
    w = np.full(P["n"], wv, np.int32)
    for flag in (0x08, 0x00):
        _check(emu, oracle, P, w, "map-ont", flag, G)


def test_emu_ksw_lead64_golden(emu):
    """the walk kernels detect a step into the AVX-512 build's lead-in cells and ksw_lead64_pair redoes those pairs:
    reference vectors (ksw_extd2_avx512) on which the SSE build answers differently"""
    from helpers import lead64_golden_cases
    cases = lead64_golden_cases()
    redone = 0
    for i, (q, t, sc, flag, w, differ, ez_exp, cig_exp) in enumerate(cases):
        P = synth.pack_pairs([q], [t])
        res, cig = emu.ksw_batch(P, np.array([w], np.int32), synth.score_matrix(sc["a"], sc["b"]), sc, flag, (4, 8, 32)[i % 3])
        assert [int(res[0][f]) for f in EXTZ_FIELDS] == ez_exp, "case %d" % i
        assert np.array_equal(cig[0][:ez_exp[EXTZ_FIELDS.index("n_cigar")]], cig_exp), "case %d" % i
        assert not differ or emu.last_lead64 == 1
        redone += emu.last_lead64
    assert redone >= 40


def _tie_pairs(n, seed):
    """low-complexity pairs (short tandem repeats, homopolymers): many equal scores along an anti-diagonal,
    which is what the reference's 4-lane row-maximum tie order is sensitive to"""
    rng = np.random.default_rng(seed)
    qs, ts = [], []
    for i in range(n):
        unit = rng.integers(0, 4, int(rng.integers(1, 4)), dtype=np.uint8)
        tl, ql = int(rng.integers(8, 150)), int(rng.integers(8, 150))
        t = np.resize(unit, tl).copy()
        q = np.resize(unit, ql).copy()
        for a in (t, q):
            k = int(rng.integers(0, 4))
            a[rng.integers(0, len(a), k)] = rng.integers(0, 4, k, dtype=np.uint8)
        qs.append(q)
        ts.append(t)
    return synth.pack_pairs(qs, ts)


@pytest.mark.parametrize("flag", [0x00, 0x40, 0x02])
def test_emu_ksw_exact_max_ties(emu, oracle, flag):
    P = _tie_pairs(16, seed=21 + flag)
    rng = np.random.default_rng(flag)
    for G, scn in ((4, "sr"), (8, "map-ont")):
        w = rng.choice([-1, 7, 20, 64, 200], P["n"]).astype(np.int32)
        _check(emu, oracle, P, w, scn, flag, G)


def test_emu_ksw_literal_row_max(emu, oracle):
    """the fallback scan taken when a row does not fit the relative 16-bit keys (forced here)"""
    P = synth.ragged_pairs(8, seed=31, max_len=120)
    T = _tie_pairs(8, seed=32)
    rng = np.random.default_rng(5)
    for Q in (P, T):
        w = rng.choice([-1, 9, 33, 100], Q["n"]).astype(np.int32)
        _check(emu, oracle, Q, w, "sr", 0x00, 4, force_literal_max=True)


@pytest.mark.parametrize("G", [64, 128])
def test_emu_ksw_block_per_pair(emu, oracle, G):
    """G > 32: one pair per thread block (the widest bands); block-level barriers and reductions replace the warp ones"""
    P = synth.ragged_pairs(6, seed=41 + G, max_len=900)
    rng = np.random.default_rng(G)
    w = rng.choice([-1, 64, 300, 700], P["n"]).astype(np.int32)
    for flag, scn in ((0x08, "map-ont"), (0x00, "map-hifi"), (0x18, "sr")):
        _check(emu, oracle, P, w, scn, flag, G)


def test_emu_ksw_microbench_shape(emu, oracle):
    P = synth.ksw_pairs(6, 150, 200, 0.05, seed=3, n_every=2)
    w = np.full(P["n"], 150, np.int32)
    for flag in (0x08, 0x00, 0x40):
        _check(emu, oracle, P, w, "sr", flag, 8)


def test_emu_sketch(emu, oracle):
    rng = np.random.default_rng(3)
    cfgs = [(21, 11), (19, 19), (15, 10), (28, 8), (17, 30), (11, 9), (12, 5)]
    pats = ["10", "110", "1110", "100", "11", "101001", "1"]
    for it in range(14):
        k, w = cfgs[it % 7]
        Z = pats[it % 6]
        small = it % 2
        if small:
            tp = 256 - (2 * w + k - 3) - (w - 1)  # emit positions of the one-warp tile (gd_sketch.cuh: sk_tile_emit)
            lens = [int(rng.integers(len(Z), tp * len(Z) // Z.count("1"))) for _ in range(10)]
        else:
            lens = [int(x) for x in rng.choice([40, 150, 1000, 5000, 9000], 5)]
        seqs = []
        for n in lens:
            c = rng.integers(0, 4, n)
            if it % 3 == 1:
                c[rng.random(n) < 0.01] = 4
            if it % 3 == 2:
                a = int(rng.integers(0, n))
                c[a:a + int(rng.integers(1, 40))] = 4
            seqs.append(bytes(synth.ACGTN[c]))
        shifts = [int(rng.integers(0, len(Z))) for _ in seqs]
        got = emu.sketch_jobs(seqs, shifts, list(range(len(seqs))), w, k, Z, small)
        for i, (s, sh) in enumerate(zip(seqs, shifts)):
            exp, _ = oracle.mm_sketch3(s, w, k, i, Z, sh, 0)
            assert np.array_equal(exp, got[i]), (it, i)


@pytest.mark.parametrize("ver", [3, 2])
def test_emu_sketch_geometry(emu, oracle, ver):
    """Window / k-mer geometries at the edges of the tile body's cases: w = 9, 16, 17, 24, 25 (how many whole 8-position
    chunks lie inside a window), 2k = 32 / 34 (32-bit against two-word hashing), k = 28, tiny k, multi-tile jobs whose
    tile seams fall at every phase of the pattern, N runs across seams."""
    emu.lib.emu_sketch_version(ver)
    try:
        rng = np.random.default_rng(11)
        cfgs = [(16, 9), (17, 16), (28, 17), (9, 24), (21, 25), (4, 33), (15, 50), (8, 12), (16, 3), (17, 1)]
        pats = ["10", "1", "110", "100", "10110", "10"]
        for it, (k, w) in enumerate(cfgs):
            Z = pats[it % len(pats)]
            for small in (0, 1):
                if small:
                    tp = 256 - (2 * w + k - 3) - (w - 1)
                    if tp < 8:
                        continue
                    lens = [int(rng.integers(len(Z), max(len(Z) + 1, tp * len(Z) // Z.count("1")))) for _ in range(6)]
                else:
                    lens = [int(x) for x in rng.choice([300, 2100, 4500, 7000], 3)]
                seqs = []
                for n in lens:
                    c = rng.integers(0, 4, n)
                    if k >= 12 and it % 2 == 1:
                        for _ in range(3):
                            a = int(rng.integers(0, n))
                            c[a:a + int(rng.integers(1, 30))] = 4
                    seqs.append(bytes(synth.ACGTN[c]))
                shifts = [int(rng.integers(0, len(Z))) for _ in seqs]
                got = emu.sketch_jobs(seqs, shifts, list(range(len(seqs))), w, k, Z, small)
                for i, (sq, sh) in enumerate(zip(seqs, shifts)):
                    exp, _ = oracle.mm_sketch3(sq, w, k, i, Z, sh, 0)
                    assert np.array_equal(exp, got[i]), (ver, k, w, Z, small, i)
    finally:
        emu.lib.emu_sketch_version(3)


def test_emu_sketch_packed_tiles(emu, oracle):
    """Several whole jobs per tile (the short-read configuration: fixed-stride output, one N slot between the jobs): every job's
    list equals mm_sketch3 of that job alone -- jobs of mixed lengths, jobs too short to emit (they get no slots), Ns next to
    the seams, one / two / four warps per tile, a last tile that is not full."""
    rng = np.random.default_rng(23)
    cases = [(21, 11, "10", 6, 64, 150), (21, 11, "10", 3, 32, 150), (15, 10, "10", 12, 128, 150), (19, 9, "110", 4, 64, 120),
             (12, 5, "10", 6, 32, 60), (28, 17, "1", 2, 64, 100), (17, 12, "101", 5, 128, 250)]
    for k, w, Z, pack, threads, maxlen in cases:
        W, ones = len(Z), Z.count("1")
        seg = maxlen // W * ones + ones + 1
        assert pack * seg <= threads * 8 - (w - 1), (pack, seg)
        seqs, shifts = [], []
        for i in range(2 * pack + 3):
            n = int(rng.integers(max(W, 8), maxlen + 1)) if i % 4 else maxlen
            if i % 5 == 3:
                n = int(rng.integers(W, 20))  # cannot emit
            c = rng.integers(0, 4, n)
            if i % 3 == 1 and k >= 12:
                c[[0, n - 1]] = 4
                a = int(rng.integers(0, n))
                c[a:a + int(rng.integers(1, 6))] = 4
            seqs.append(bytes(synth.ACGTN[c]))
            shifts.append(int(rng.integers(0, W)))
        got = emu.sketch_packed(seqs, shifts, list(range(len(seqs))), w, k, Z, pack, threads)
        for i, (sq, sh) in enumerate(zip(seqs, shifts)):
            exp, _ = oracle.mm_sketch3(sq, w, k, i, Z, sh, 0)
            assert np.array_equal(exp, got[i]), (k, w, Z, pack, threads, i)
    # one job per tile through the same entry (tickets in fixed-stride mode)
    got = emu.sketch_packed(seqs, shifts, list(range(len(seqs))), w, k, Z, 0, 32)
    for i, (sq, sh) in enumerate(zip(seqs, shifts)):
        exp, _ = oracle.mm_sketch3(sq, w, k, i, Z, sh, 0)
        assert np.array_equal(exp, got[i]), i


def test_emu_sketch_concurrent_blocks(emu, oracle):
    """The protocol BETWEEN blocks of the dense sketch kernel -- ordering tickets, a tile's count published before its
    offset is known, records parked in shared memory, look-back and copy-out inside the block's next tile -- with all
    blocks resident and scheduled in pseudo-random interleavings: the output must be the reference's, in order, whatever
    the interleaving (and no interleaving may dead-lock).  Also the one-warp dense tiles, the packed tiles and the v2 body."""
    rng = np.random.default_rng(5)
    seqs = [bytes(synth.ACGTN[rng.integers(0, 4, n)]) for n in (30000, 9000, 40, 17000, 4200)]
    exp = {}
    try:
        for it, (k, w, Z, grid) in enumerate([(21, 11, "10", 4), (15, 10, "10", 6), (19, 19, "110", 3), (12, 3, "1", 5)]):
            shifts = [int(rng.integers(0, len(Z))) for _ in seqs]
            want = [oracle.mm_sketch3(s, w, k, i, Z, sh, 0)[0] for i, (s, sh) in enumerate(zip(seqs, shifts))]
            for seed in (1, 2):
                emu.sketch_concurrency(1000 * it + seed)
                got = emu.sketch_jobs(seqs, shifts, list(range(len(seqs))), w, k, Z, 0, grid=grid)
                for i in range(len(seqs)):
                    assert np.array_equal(want[i], got[i]), (k, w, Z, grid, seed, i)
        # one-warp tiles with dense output (look-back over one-tile jobs), packed tiles, and the previous body
        short = [bytes(synth.ACGTN[rng.integers(0, 4, int(n))]) for n in rng.integers(30, 300, 24)]
        sh = [int(rng.integers(0, 2)) for _ in short]
        want = [oracle.mm_sketch3(s, 11, 21, i, "10", x, 0)[0] for i, (s, x) in enumerate(zip(short, sh))]
        for seed in (7, 8):
            emu.sketch_concurrency(seed)
            got = emu.sketch_jobs(short, sh, list(range(len(short))), 11, 21, "10", 1, grid=5)
            assert all(np.array_equal(a, b) for a, b in zip(want, got)), seed
            got = emu.sketch_packed(short, sh, list(range(len(short))), 11, 21, "10", 3, 64, grid=4)
            assert all(np.array_equal(a, b) for a, b in zip(want, got)), seed
        emu.lib.emu_sketch_version(2)
        emu.sketch_concurrency(11)
        got = emu.sketch_jobs(seqs[:3], [0, 1, 0], [0, 1, 2], 11, 21, "10", 0, grid=3)
        for i in range(3):
            assert np.array_equal(oracle.mm_sketch3(seqs[i], 11, 21, i, "10", [0, 1, 0][i], 0)[0], got[i])
    finally:
        emu.lib.emu_sketch_version(3)
        emu.sketch_concurrency(0)


def test_emu_ksw_shuffled_schedules(emu, oracle):
    """The DP kernel with blocks AND the threads inside a block scheduled in pseudo-random orders between barriers / warp
    collectives (simt_emu.h: launch_concurrent): a missing __syncwarp / __syncthreads around the shared-memory ring would give
    a wrong result for some order.  Warp gangs of 4 / 8 / 32 lanes and the block-per-pair gangs."""
    emu.lib.emu_set_shuffle.argtypes = [C.c_uint]
    try:
        for seed, (G, flag, mx) in enumerate([(4, 0x00, 160), (4, 0x08, 160), (8, 0x40, 200), (32, 0x00, 260), (64, 0x08, 420), (128, 0x00, 420)]):
            emu.lib.emu_set_shuffle(101 + seed)
            P = synth.ragged_pairs(6, seed=40 + seed, max_len=mx)
            w = np.minimum(np.maximum(P["qlen"], P["tlen"]), 150).astype(np.int32)
            _check(emu, oracle, P, w, "sr", flag, G)
    finally:
        emu.lib.emu_set_shuffle(0)


def _asan_run(tmp_path, driver, kernel_src, timeout):
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "asan_driver")
    csrc = os.path.join(os.path.dirname(here), "genome-on-diet_b200", "csrc")
    cmd = ["g++", "-O1", "-fsanitize=address", "-fno-omit-frame-pointer", "-std=c++17", "-I", os.path.join(here, "emu"), "-I", csrc,
           os.path.join(here, "emu", driver), os.path.join(here, "emu", kernel_src), "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("this g++ has no AddressSanitizer runtime")
    assert b.returncode == 0, b.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0:detect_leaks=0")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=timeout)
    assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stderr, (r.stdout[-1000:], r.stderr[-3000:])
    return r.stdout


def test_emu_sketch_address_sanitizer(tmp_path):
    """The sketch kernels' device code under AddressSanitizer (exact-size heap buffers for the sequences, outputs, job and
    status arrays and the blocks' shared memory): dense, one-warp and packed tiles over random geometries, sequential and
    interleaved blocks.  Out-of-bounds shared-memory reads do not fault on a GPU; here they abort."""
    out = _asan_run(tmp_path, "asan_sketch_driver.cpp", "emu_sketch.cpp", 600)
    assert out.count("packed T") >= 12


def test_emu_ksw_address_sanitizer(tmp_path):
    """The DP, traceback and lead-in kernels' device code under AddressSanitizer: random pairs (1..420 bases), bands 0..460,
    ten flag sets, every gang size from 4 lanes to a 128-thread block per pair."""
    out = _asan_run(tmp_path, "asan_ksw_driver.cpp", "emu_ksw.cpp", 900)
    assert out.count(" rc 0 ") == 10
