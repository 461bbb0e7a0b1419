"""Host logic after the DP (mm_update_extra / mm_fix_cigar, candidate filter, mm_set_sam_params,
mm_write_sam3) in libgdiet_cuda.so's plain-C++ part, fed with the ORACLE's candidates so that it runs without a GPU,
against the SAM text of the unmodified reference program (golden fixture + live runs where oracle/_ref travelled)."""
import gzip
import os

import numpy as np
import pytest

import gdiet_b200 as gd
import maplib
from oraclelib import cpu_has_avx512

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_candidates(M, contigs, reads, o, Z="10"):
    mi = M.index_build(contigs, 11, 21, Z)
    cand_off, cands, cigs, base = [0], [], [], 0
    for r in reads:
        c, cig, _ = M.map_read(mi, r, o)
        c = c.copy()
        used = int(sum(max(int(x), 0) for x in c["n_cigar"]))
        c["cigar_off"] += base
        cands.append(c)
        cigs.append(cig[:used])
        base += used
        cand_off.append(cand_off[-1] + len(c))
    M.lib.gdo_index_destroy(mi)
    return (np.array(cand_off, np.int64), np.concatenate(cands) if cands else np.zeros(0, maplib.CAND_DTYPE),
            np.concatenate(cigs) if cigs else np.zeros(0, np.uint32))


def our_sam(contigs, reads, cand_off, cand, cig, post):
    n, L = reads.shape
    names = ["r%d" % i for i in range(n)]
    seq_names = ["chr%d" % (i + 1) for i in range(len(contigs))]
    off = np.arange(n, dtype=np.int64) * L
    lens = np.full(n, L, np.int32)
    qual = np.full(n * L, ord("I"), np.uint8)
    hdr = gd.sam_header(seq_names, [len(c) for c in contigs])
    body = gd.sr_sam_batch(names, off, lens, np.ascontiguousarray(reads.reshape(-1)), qual, cand_off, cand, cig, seq_names, contigs, post)
    return (hdr + body).decode().splitlines()


def test_sam_parts_equal_the_concatenated_text():
    g = np.load(os.path.join(GOLDEN, "map_sr.npz"))
    contigs, reads = maplib.make_dataset(seed=int(g["seed"]), n_reads=int(g["n_reads"]))
    o = maplib.sr_opt(min_cnt=float(g["min_cnt"]), rec_frac=float(g["rec_frac"]))
    cand_off, cand, cig = oracle_candidates(maplib.MapOracle(), contigs, reads, o)
    n, L = reads.shape
    names = ["r%d" % i for i in range(n)]
    args = (names, np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int32), np.ascontiguousarray(reads.reshape(-1)),
            np.full(n * L, ord("I"), np.uint8), cand_off, cand, cig, ["chr1", "chr2", "chr3"], contigs, gd.sr_post_options(n_threads=5))
    whole = gd.sr_sam_batch(*args)
    parts = gd.sr_sam_batch(*args, parts=True)
    assert parts.count == 5 and parts.bytes() == whole
    parts.free()


def strip_pg(text):
    return [l for l in text.splitlines() if not l.startswith("@PG")]


def test_sam_golden():
    g = np.load(os.path.join(GOLDEN, "map_sr.npz"))
    with gzip.open(os.path.join(GOLDEN, "map_sr.sam.gz"), "rt") as f:
        want = strip_pg(f.read())
    contigs, reads = maplib.make_dataset(seed=int(g["seed"]), n_reads=int(g["n_reads"]))
    o = maplib.sr_opt(min_cnt=float(g["min_cnt"]), rec_frac=float(g["rec_frac"]))
    M = maplib.MapOracle()
    cand_off, cand, cig = oracle_candidates(M, contigs, reads, o)
    got = our_sam(contigs, reads, cand_off, cand, cig, gd.sr_post_options())
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a == b


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_sr and AVX-512")
@pytest.mark.parametrize("seed,read_len,extra,okw,pkw", [
    (41, 150, ["-r", "0.05,150,200"], {}, {}),
    (42, 150, ["-r", "0.05,150,200", "-n", "0.1", "--secondary=yes"], dict(min_cnt=0.1), dict(no_print_2nd=0)),
    (43, 400, ["-n", "0.2"], dict(min_cnt=0.2, bw_min=500, bw_max=1500), {}),
    (44, 100, ["-r", "0.1,20,50", "-n", "0.1", "-Y"], dict(min_cnt=0.1, bw_frac=0.1, bw_min=20, bw_max=50), dict(softclip=1)),
])
def test_sam_matches_reference_program(seed, read_len, extra, okw, pkw):
    contigs, reads = maplib.make_dataset(seed=seed, read_len=read_len, n_reads=1500)
    o = maplib.sr_opt(qlen=read_len, **okw)
    sam, _ = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=extra), trace=False, threads=2)
    M = maplib.MapOracle()
    cand_off, cand, cig = oracle_candidates(M, contigs, reads, o)
    got = our_sam(contigs, reads, cand_off, cand, cig, gd.sr_post_options(**pkw))
    want = strip_pg(sam)
    assert len(got) == len(want)
    bad = [(a, b) for a, b in zip(got, want) if a != b]
    assert not bad, "%d differing SAM lines, first:\n%s\n%s" % (len(bad), bad[0][0], bad[0][1])


# ---- long-read tree -----------------------------------------------------------------------------------------------
def lr_oracle_candidates(M, mi, reads, o):
    cand_off, cands, cigs, base = [0], [], [], 0
    for r in reads:
        c, cig, _ = M.lr_map_read(mi, r, o)
        c = c.copy()
        used = int(sum(max(int(x), 0) for x in c["n_cigar"]))
        c["cigar_off"] += base
        cands.append(c)
        cigs.append(cig[:used])
        base += used
        cand_off.append(cand_off[-1] + len(c))
    return np.array(cand_off, np.int64), np.concatenate(cands), np.concatenate(cigs) if cigs else np.zeros(0, np.uint32)


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_lr and AVX-512")
@pytest.mark.parametrize("case", [
    (1, "map-hifi", 19, 19, 1000, 15000, 0.005, 0.005, 30, [], {}),
    (3, "map-hifi", 19, 19, 400, 6000, 0.005, 0.005, 60, ["--vt_nb_loc=2"], dict(vt_nb_loc=2)),
    (13, "map-ont", 15, 10, 800, 8000, 0.02, 0.03, 60,
     ["--vt_dis=1000", "--vt_nb_loc=3", "--vt_df1=0.007", "--vt_df2=0.007", "--max_min_gap=4000", "--vt_f=0.04", "--vt_cov", "0.3",
      "--sort=merge", "--frag=no"], dict(vt_dis=1000, vt_df1=0.007, vt_df2=0.007, vt_f=0.04, vt_cov=0.3)),
])
def test_lr_sam_matches_reference_program(case):
    """gd_lr_sam_batch: every read -- including those whose chained candidates go through the concatenate_cigars
    restatement -- gets exactly the reference program's SAM records."""
    from test_oracle_map_lr_vs_ref import lr_setup
    M = maplib.MapOracle()
    contigs, reads, flags, mi, o = lr_setup(M, *case)
    sam, _ = maplib.run_reference(contigs, reads, flags, program=maplib.REF_LR, trace=False, threads=2)
    cand_off, cand, cig = lr_oracle_candidates(M, mi, reads, o)
    M.lib.gdo_index_destroy(mi)
    lens = np.array([len(r) for r in reads], np.int32)
    off = np.zeros(len(reads), np.int64)
    off[1:] = np.cumsum(lens[:-1].astype(np.int64))
    buf = np.concatenate(reads)
    names = ["r%d" % i for i in range(len(reads))]
    txt, sam_off, stitch = gd.lr_sam_batch(names, off, lens, buf, np.full(len(buf), ord("I"), np.uint8), cand_off, cand, cig,
                                           ["chr%d" % (i + 1) for i in range(len(contigs))], contigs, gd.lr_post_options(case[1]))
    want = {}
    for l in sam.splitlines():
        if not l.startswith("@"):
            want.setdefault(l.split("\t", 1)[0], []).append(l)
    n_chain = bad = 0
    first_bad = None
    for i, nm in enumerate(names):
        mine = txt[sam_off[i]:sam_off[i + 1]].decode().splitlines()
        c = cand[cand_off[i]:cand_off[i + 1]]
        n_chain += int(len(c) > 0 and (c["reserved"][:, 0] >= 0).any())
        if mine != want[nm]:
            bad += 1
            first_bad = first_bad or (nm, [m[:300] for m in mine], [m[:300] for m in want[nm]])
    assert bad == 0, "%d reads differ, first: %s" % (bad, first_bad)
    assert n_chain > 0 and stitch.sum() == 0      # reads with chained candidates went through concatenate_cigars


def test_sam_core_of_the_gpu_path_equals_the_host_stage():
    """csrc/gd_sam_core.h -- what the GPU runs with one thread per read (gd_sr_map_sam_batch) -- executed on the host against
    gd_sr_sam_batch on oracle candidates: every option set of the short-read tests, N-holding and ragged reads, reads with
    several candidates (repeat), soft clips, secondary records printed"""
    from emulib import Emu
    E = Emu()
    assert E.lib.emu_sam_check_fixed4(700) == 0  # "%.4f" by exact integer rounding == printf for every mlen/den up to 700
    M = maplib.MapOracle()
    rng = np.random.default_rng(9)
    for seed, okw, pkw, ragged in ((1, {}, {}, False), (2, dict(min_cnt=0.2, rec_frac=0.1), dict(no_print_2nd=0, best_n=5), False),
                                   (7, dict(min_cnt=0.2, bw_min=500, bw_max=1500, af_max_loc=2), dict(softclip=1), True),
                                   (8, dict(min_cnt=0.2, mid_occ=2, max_max_occ=3, occ_dist=40), dict(sam_hit_only=1, no_print_2nd=0), False)):
        contigs, reads = maplib.make_dataset(seed=seed, n_reads=1200)
        rl = [r[:int(rng.integers(40, 151))].copy() for r in reads] if ragged else list(reads)
        o = maplib.sr_opt(**okw)
        cand_off, cand, cig = oracle_candidates(M, contigs, rl, o)
        lens = np.array([len(r) for r in rl], np.int32)
        off = np.zeros(len(rl), np.int64)
        off[1:] = np.cumsum(lens[:-1].astype(np.int64))
        buf = np.concatenate(rl)
        qual = (33 + (np.arange(len(buf)) % 40)).astype(np.uint8)
        names = ["read_%d" % i for i in range(len(rl))]
        seq_names = ["chr%d" % (i + 1) for i in range(len(contigs))]
        post = gd.sr_post_options(n_threads=3, **pkw)
        for q in (qual, None):
            want = gd.sr_sam_batch(names, off, lens, buf, q, cand_off, cand, cig, seq_names, contigs, post)
            got = E.sam_batch(names, off, lens, buf, q, cand_off, cand, cig, seq_names, contigs, post)
            assert got == want, "seed %d" % seed
