"""GPU parity of the device-resident index and the short-read mapping stage against
oracle/gd_oracle_map.c, the reference call trace (when oracle/_ref/GDiet_avx_sr travelled) and the golden fixtures."""
import ctypes
import os

import numpy as np
import pytest

import maplib
from oraclelib import cpu_has_avx512

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def M():
    return maplib.MapOracle()


def flat_reads(reads):
    n, L = reads.shape
    return np.arange(n, dtype=np.int64) * L, np.full(n, L, np.int32), np.ascontiguousarray(reads.reshape(-1))


def test_index_matches_oracle(ctx, M):
    contigs, _ = maplib.make_dataset(seed=1, n_reads=1)
    for Z in ("10", "110"):
        idx = ctx.index_build(contigs, 11, 21, Z)
        mi = M.index_build(contigs, 11, 21, Z)
        keys, counts, pos = M.index_arrays(mi)
        dk, dc, dp, S = idx.export()
        assert np.array_equal(dk, keys) and np.array_equal(dc, counts) and np.array_equal(dp, pos)
        # mi->S: 4-bit codes, 8 per word, contigs back to back (index.c:351-356)
        codes = np.concatenate([np.searchsorted(np.frombuffer(b"ACGT", np.uint8), c) for c in contigs]).astype(np.uint8)
        nib = ((S[:, None] >> (4 * np.arange(8, dtype=np.uint32))[None, :]) & 0xf).reshape(-1)[: len(codes)]
        assert np.array_equal(nib, codes)
        # mm_idx_get: present and absent minimizers
        rng = np.random.default_rng(0)
        probe = np.concatenate([keys[rng.integers(0, len(keys), 2000)], rng.integers(0, 1 << 42, 2000, dtype=np.uint64)])
        cnt, first = idx.get(probe)
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        look = {int(k): (int(c), int(s)) for k, c, s in zip(keys, counts, starts)}
        for m, c, f in zip(probe, cnt, first):
            e = look.get(int(m))
            assert (int(c), int(f)) == (e if e else (0, -1))
        for f in (2e-4, 0.01, 0.5):
            assert idx.cal_max_occ(f) == M.lib.gdo_index_cal_max_occ(mi, f)
        idx.close()
        M.lib.gdo_index_destroy(mi)


CASES = [
    (1, "10", 150, {}),
    (2, "10", 150, dict(min_cnt=0.2, rec_frac=0.1)),
    (3, "110", 150, dict(min_cnt=0.3)),
    (4, "10", 400, dict(min_cnt=0.2, bw_min=500, bw_max=1500)),
    (5, "10", 100, dict(min_cnt=0.1, bw_frac=0.1, bw_min=20, bw_max=50)),
    (7, "10", 300, dict(min_cnt=0.2, bw_min=500, bw_max=1500, af_max_loc=2)),
    (8, "10", 150, dict(min_cnt=0.2, mid_occ=2, max_max_occ=3, occ_dist=40)),   # mm_seed_select on the planted repeat
    (9, "10", 150, dict(min_cnt=0.2, mid_occ=2, max_max_occ=2, occ_dist=0)),    # plain max_occ filter
    (10, "10", 150, dict(min_cnt=0.1, for_only=1)),
    (11, "10", 150, dict(min_cnt=0.1, rev_only=1)),
]


@pytest.mark.parametrize("seed,Z,read_len,okw", CASES)
def test_sr_map_matches_oracle(ctx, M, seed, Z, read_len, okw):
    contigs, reads = maplib.make_dataset(seed=seed, read_len=read_len, n_reads=1500)
    o = maplib.sr_opt(Z=Z, qlen=read_len, **okw)
    idx = ctx.index_build(contigs, 11, 21, Z)
    mi = M.index_build(contigs, 11, 21, Z)
    off, lens, buf = flat_reads(reads)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    total = 0
    for i, r in enumerate(reads):
        oc, ocig, dbg = M.map_read(mi, r, o)
        mine = cand[coff[i]:coff[i + 1]]
        assert len(mine) == len(oc), "read %d: %d candidates, oracle %d (%s)" % (i, len(mine), len(oc), dbg)
        for j, (a, b) in enumerate(zip(mine, oc)):
            for f in ("rid", "rs", "re", "qs", "qe", "rev", "votes", "first_q", "last_q", "exact", "score", "n_cigar"):
                assert int(a[f]) == int(b[f]), "read %d cand %d field %s: %d vs oracle %d (%s)" % (i, j, f, a[f], b[f], dbg)
            ga = cig[int(a["cigar_off"]):int(a["cigar_off"]) + max(int(a["n_cigar"]), 0)]
            gb = ocig[int(b["cigar_off"]):int(b["cigar_off"]) + max(int(b["n_cigar"]), 0)]
            assert np.array_equal(ga, gb), "read %d cand %d cigar" % (i, j)
        total += len(mine)
    assert total > 700
    idx.close()
    M.lib.gdo_index_destroy(mi)


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_sr and AVX-512")
def test_sr_map_matches_reference_program(ctx):
    """The device stage against the call trace of the unmodified reference program, config-1 flags."""
    contigs, reads = maplib.make_dataset(seed=21, n_reads=4000, contig_lens=(1000000, 400000, 100000))
    o = maplib.sr_opt()
    _, tr = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=["-r", "0.05,150,200"]))
    idx = ctx.index_build(contigs, 11, 21, "10")
    off, lens, buf = flat_reads(reads)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    for i, t in enumerate(tr):
        maplib.cands_equal_trace(cand[coff[i]:coff[i + 1]], cig, t["cands"], "read %d" % i)
    idx.close()


def test_sr_map_golden(ctx):
    """Committed vectors generated from the reference program by tests/golden/make_golden_map.py."""
    g = np.load(os.path.join(GOLDEN, "map_sr.npz"))
    contigs, reads = maplib.make_dataset(seed=int(g["seed"]), n_reads=int(g["n_reads"]))
    o = maplib.sr_opt(min_cnt=float(g["min_cnt"]), rec_frac=float(g["rec_frac"]))
    idx = ctx.index_build(contigs, 11, 21, "10")
    off, lens, buf = flat_reads(reads)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    assert np.array_equal(coff, g["cand_off"])
    for f in ("rid", "rs", "re", "qs", "qe", "rev", "exact", "score", "n_cigar"):
        assert np.array_equal(cand[f], g[f]), f
    mine = np.concatenate([cig[int(c["cigar_off"]):int(c["cigar_off"]) + max(int(c["n_cigar"]), 0)] for c in cand])
    assert np.array_equal(mine, g["cigar"])
    idx.close()


def test_sr_map_empty_and_unmappable(ctx):
    contigs, _ = maplib.make_dataset(seed=1, n_reads=1)
    idx = ctx.index_build(contigs, 11, 21, "10")
    o = maplib.sr_opt()
    coff, cand, cig = ctx.sr_map_batch(idx, np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(1, np.uint8), o)
    assert len(cand) == 0 and coff[0] == 0
    junk = np.frombuffer(b"ACGT" * 40, np.uint8)[None, :150].repeat(7, 0).copy()
    junk[3] = np.frombuffer(b"N" * 150, np.uint8)
    off, lens, buf = flat_reads(junk)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    assert len(cand) == 0 and np.all(coff == 0)
    idx.close()


def test_sam_end_to_end_golden(ctx):
    """Device mapping stage + host post-processing: the SAM text equals the reference program's, byte for byte."""
    import gzip
    import gdiet_b200 as gd
    g = np.load(os.path.join(GOLDEN, "map_sr.npz"))
    with gzip.open(os.path.join(GOLDEN, "map_sr.sam.gz"), "rt") as f:
        want = [l for l in f.read().splitlines() if not l.startswith("@")]
    contigs, reads = maplib.make_dataset(seed=int(g["seed"]), n_reads=int(g["n_reads"]))
    o = maplib.sr_opt(min_cnt=float(g["min_cnt"]), rec_frac=float(g["rec_frac"]))
    idx = ctx.index_build(contigs, 11, 21, "10")
    off, lens, buf = flat_reads(reads)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    names = ["r%d" % i for i in range(len(reads))]
    qual = np.full(len(buf), ord("I"), np.uint8)
    sam = gd.sr_sam_batch(names, off, lens, buf, qual, coff, cand, cig, ["chr1", "chr2", "chr3"], contigs, gd.sr_post_options())
    assert sam.decode().splitlines() == want
    idx.close()


# ---- long-read tree ---------------------------------------------------------------------------------------------
def flat_ragged(reads):
    lens = np.array([len(r) for r in reads], np.int32)
    off = np.zeros(len(reads), np.int64)
    off[1:] = np.cumsum(lens[:-1].astype(np.int64))
    return off, lens, np.concatenate(reads)


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_lr and AVX-512")
@pytest.mark.parametrize("case", [
    (1, "map-hifi", 19, 19, 1000, 15000, 0.005, 0.005, 120, [], {}),
    (3, "map-hifi", 19, 19, 400, 6000, 0.005, 0.005, 200, ["--vt_nb_loc=2"], dict(vt_nb_loc=2)),
    (2, "map-ont", 15, 10, 1300, 20000, 0.03, 0.05, 100,
     ["--vt_dis=1000", "--vt_nb_loc=3", "--vt_df1=0.007", "--vt_df2=0.007", "--max_min_gap=4000", "--vt_f=0.04", "--vt_cov", "0.3",
      "--sort=merge", "--frag=no"], dict(vt_dis=1000, vt_df1=0.007, vt_df2=0.007, vt_f=0.04, vt_cov=0.3)),
])
def test_lr_map_matches_reference_program(ctx, case):
    """gd_lr_map_batch against the call trace of the unmodified long-read reference program (HiFi- and ONT-like reads
    with structural variation, chimeras and unmappable reads)."""
    import gdiet_b200 as gd
    seed, preset, k, w, bw, read_len, sub, indel, n_reads, extra, okw = case
    contigs, reads = maplib.make_long_dataset(seed=seed, read_len=read_len, sub=sub, indel=indel, n_reads=n_reads)
    flags = ["-ax", preset, "-Z", "10", "-W", "2", "-k", str(k), "-w", str(w), "-r", str(bw)] + list(extra)
    _, tr = maplib.run_reference(contigs, reads, flags, program=maplib.REF_LR, threads=1)
    idx = ctx.index_build(contigs, w, k, "10")
    lo, hi = (50, 500) if preset == "map-hifi" else (10, 1000000)
    mid = min(max(idx.cal_max_occ(2e-4), lo), hi)
    o = gd.lr_options(preset, bw=bw, mid_occ=mid, **okw)
    off, lens, buf = flat_ragged(reads)
    coff, cand, cig = ctx.lr_map_batch(idx, off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=64 * len(reads) * 256)
    n_cand = n_chain = 0
    for i, t in enumerate(tr):
        mine = cand[coff[i]:coff[i + 1]]
        maplib.lr_cands_equal_trace(mine, cig, t["cands"], "read %d" % i)
        n_cand += len(mine)
        n_chain += int((mine["reserved"][:, 0] >= 0).sum()) if len(mine) else 0
    assert n_cand >= len(reads) // 3 and n_chain > 0
    idx.close()


def test_lr_map_matches_oracle(ctx, M):
    """... and against oracle/gd_oracle_map.c including vt_t::next / concat (no reference program needed)."""
    import gdiet_b200 as gd
    contigs, reads = maplib.make_long_dataset(seed=5, read_len=5000, sub=0.005, indel=0.005, n_reads=40)
    idx = ctx.index_build(contigs, 19, 19, "10")
    mi = M.index_build(contigs, 19, 19, "10")
    o = gd.lr_options("map-hifi", bw=300, mid_occ=50)
    off, lens, buf = flat_ragged(reads)
    coff, cand, cig = ctx.lr_map_batch(idx, off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=len(reads) * 4096)
    for i, r in enumerate(reads):
        oc, ocig, dbg = M.lr_map_read(mi, r, o)
        mine = cand[coff[i]:coff[i + 1]]
        assert len(mine) == len(oc), "read %d: %d vs oracle %d (%s)" % (i, len(mine), len(oc), dbg)
        for j, (a, b) in enumerate(zip(mine, oc)):
            for f in ("rid", "rs", "re", "qs", "qe", "rev", "votes", "first_q", "last_q", "exact", "score", "n_cigar"):
                assert int(a[f]) == int(b[f]), "read %d cand %d field %s: %d vs oracle %d" % (i, j, f, a[f], b[f])
            assert tuple(a["reserved"][:2]) == tuple(b["reserved"][:2]), "read %d cand %d next/concat" % (i, j)
            ga = cig[int(a["cigar_off"]):int(a["cigar_off"]) + max(int(a["n_cigar"]), 0)]
            gb = ocig[int(b["cigar_off"]):int(b["cigar_off"]) + max(int(b["n_cigar"]), 0)]
            assert np.array_equal(ga, gb), "read %d cand %d cigar" % (i, j)
    idx.close()
    M.lib.gdo_index_destroy(mi)


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_sr and AVX-512")
def test_device_index_dumps_the_reference_mmi(ctx):
    """Row F4: the index built on the device, written with gd_mmi_write, is byte-identical to `GDiet_avx -d`."""
    import hashlib
    import tempfile
    import gdiet_b200 as gd
    from test_mmi import reference_mmi
    contigs, _ = maplib.make_dataset(seed=12, contig_lens=(700000, 250000, 50001), n_reads=1)
    tmp = tempfile.mkdtemp(prefix="gdmmi_")
    want = reference_mmi(contigs, 21, 11, "10", tmp)
    idx = ctx.index_build(contigs, 11, 21, "10")
    keys, counts, pos, S = idx.export()
    out = os.path.join(tmp, "ours.mmi")
    gd.mmi_write(out, 11, 21, ["chr1", "chr2", "chr3"], [len(c) for c in contigs], keys, counts, pos, S)
    got = open(out, "rb").read()
    assert len(got) == len(want) and hashlib.md5(got).hexdigest() == hashlib.md5(want).hexdigest()
    idx.close()


def tandem_dataset(seed=7, unit=150, copies=90, n_reads=24, read_len=15000):
    """Reads that cross a long tandem repeat: one minimizer value then occurs more than mid_occ times INSIDE a read
    (mm_seed_mz_flt removes it, seed.c:5-29) and more than mid_occ times in the index (mm_seed_select, seed.c:67-113)."""
    rng = np.random.default_rng(seed)
    g = synth_genome(900000, seed)
    u = g[1000:1000 + unit].copy()
    st = 400000
    g[st:st + unit * copies] = np.tile(u, copies)
    lut = np.zeros(256, np.uint8)
    lut[maplib.synth.ACGTN] = np.arange(5)
    reads = []
    for i in range(n_reads):
        a = st - int(rng.integers(200, read_len - unit * copies - 200)) if i % 3 else int(rng.integers(0, len(g) - read_len))
        codes = maplib.synth.mutate_codes(rng, lut[g[a:a + read_len + 400]], 0.002, sub=0.5, dele=0.25)[:read_len]
        if rng.random() < 0.5:
            codes = (3 - codes)[::-1]
        reads.append(maplib.synth.ACGTN[codes])
    return [g], reads


def synth_genome(n, seed):
    return maplib.synth.random_genome(n, seed=1000 + seed)


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_lr and AVX-512")
def test_lr_map_tandem_repeats_match_reference_program(ctx):
    import gdiet_b200 as gd
    contigs, reads = tandem_dataset()
    flags = ["-ax", "map-hifi", "-Z", "10", "-W", "2", "-k", "19", "-w", "19", "-r", "1000", "-f", "60"]  # mid_occ = 60
    _, tr = maplib.run_reference(contigs, reads, flags, program=maplib.REF_LR, threads=1)
    idx = ctx.index_build(contigs, 19, 19, "10")
    o = gd.lr_options("map-hifi", bw=1000, mid_occ=60)
    off, lens, buf = flat_ragged(reads)
    coff, cand, cig = ctx.lr_map_batch(idx, off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=64 * len(reads) * 256)
    for i, t in enumerate(tr):
        maplib.lr_cands_equal_trace(cand[coff[i]:coff[i + 1]], cig, t["cands"], "read %d" % i)
    idx.close()


def test_sr_map_ragged_lowercase_and_tiny_reads(ctx, M):
    """One batch with read lengths from 2 to 330 (some shorter than k, some longer than the 300-base switch of
    map.c:776), lower-case bases and N runs, and a band that depends on the read length (-r 0.25,20,60: map.c:624-631
    computes it per read): candidates equal the oracle's read by read."""
    rng = np.random.default_rng(77)
    contigs, base_reads = maplib.make_dataset(seed=13, read_len=330, n_reads=600)
    reads = []
    for i, r in enumerate(base_reads):
        L = int(rng.choice([2, 5, 20, 21, 40, 75, 100, 149, 150, 151, 200, 299, 300, 301, 330]))
        r = r[:L].copy()
        if i % 3 == 0:
            r = np.frombuffer(bytes(r).lower(), np.uint8).copy()
        if i % 11 == 0 and L > 30:
            r[10:14] = ord("N")
        reads.append(r)
    o = maplib.sr_opt(min_cnt=0.2, rec_frac=0.1, bw_frac=0.25, bw_min=20, bw_max=60)
    idx = ctx.index_build(contigs, 11, 21, "10")
    mi = M.index_build(contigs, 11, 21, "10")
    off, lens, buf = flat_ragged(reads)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    n_cand = 0
    for i, r in enumerate(reads):
        oc, ocig, dbg = M.map_read(mi, r, o)
        mine = cand[coff[i]:coff[i + 1]]
        assert len(mine) == len(oc), "read %d (len %d): %d candidates, oracle %d (%s)" % (i, len(r), len(mine), len(oc), dbg)
        for j, (a, b) in enumerate(zip(mine, oc)):
            for f in ("rid", "rs", "re", "qs", "qe", "rev", "votes", "exact", "score", "n_cigar"):
                assert int(a[f]) == int(b[f]), "read %d (len %d) cand %d field %s: %d vs oracle %d" % (i, len(r), j, f, a[f], b[f])
            assert np.array_equal(cig[int(a["cigar_off"]):int(a["cigar_off"]) + max(int(a["n_cigar"]), 0)],
                                  ocig[int(b["cigar_off"]):int(b["cigar_off"]) + max(int(b["n_cigar"]), 0)])
        n_cand += len(mine)
    assert n_cand > 200
    idx.close()
    M.lib.gdo_index_destroy(mi)


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_lr and AVX-512")
@pytest.mark.parametrize("preset,k,w,bw,read_len,sub,indel,extra,okw", [
    ("map-hifi", 19, 19, 600, 7000, 0.005, 0.005, [], {}),
    ("map-ont", 15, 10, 800, 9000, 0.02, 0.03,
     ["--vt_dis=1000", "--vt_nb_loc=3", "--vt_df1=0.007", "--vt_df2=0.007", "--max_min_gap=4000", "--vt_f=0.04", "--vt_cov", "0.3",
      "--sort=merge", "--frag=no"], dict(vt_dis=1000, vt_df1=0.007, vt_df2=0.007, vt_f=0.04, vt_cov=0.3)),
])
def test_lr_sam_end_to_end_matches_reference_program(ctx, preset, k, w, bw, read_len, sub, indel, extra, okw):
    """Long reads end to end: gd_lr_map_batch on the device + gd_lr_sam_batch on the host (with CIGAR stitching of the
    chained candidates) against the SAM of the unmodified reference program, on reads with structural variation."""
    import gdiet_b200 as gd
    contigs, reads = maplib.make_long_dataset(seed=61, read_len=read_len, sub=sub, indel=indel, n_reads=160, sv_frac=0.6)
    flags = ["-ax", preset, "-Z", "10", "-W", "2", "-k", str(k), "-w", str(w), "-r", str(bw)] + list(extra)
    sam, _ = maplib.run_reference(contigs, reads, flags, program=maplib.REF_LR, trace=False, threads=4)
    idx = ctx.index_build(contigs, w, k, "10")
    lo, hi = (50, 500) if preset == "map-hifi" else (10, 1000000)
    o = gd.lr_options(preset, bw=bw, mid_occ=min(max(idx.cal_max_occ(2e-4), lo), hi), **okw)
    off, lens, buf = flat_ragged(reads)
    coff, cand, cig = ctx.lr_map_batch(idx, off, lens, buf, o, cand_cap=8 * len(reads), cigar_cap=64 * len(reads) * 256)
    names = ["r%d" % i for i in range(len(reads))]
    txt, sam_off, _ = gd.lr_sam_batch(names, off, lens, buf, np.full(len(buf), ord("I"), np.uint8), coff, cand, cig,
                                      ["chr%d" % (i + 1) for i in range(len(contigs))], contigs, gd.lr_post_options(preset))
    want = [l for l in sam.splitlines() if not l.startswith("@")]
    assert txt.decode().splitlines() == want
    assert int((cand["reserved"][:, 0] >= 0).sum()) > 10   # chained candidates were stitched
    idx.close()


@pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()), reason="needs oracle/_ref/GDiet_avx_sr and AVX-512")
def test_index_loaded_from_the_reference_mmi_maps_like_the_built_one(ctx):
    """Row F4, the way back: the reference program's own `-d` dump, loaded with gd_index_load_mmi, is the same device
    index as the one built from the FASTA (same arrays, same contig names) and maps reads to the same candidates."""
    import tempfile
    from test_mmi import reference_mmi
    contigs, reads = maplib.make_dataset(seed=14, contig_lens=(500000, 250000, 50001), n_reads=800)
    tmp = tempfile.mkdtemp(prefix="gdmmi_")
    open(os.path.join(tmp, "ref.mmi"), "wb").write(reference_mmi(contigs, 21, 11, "10", tmp))
    built = ctx.index_build(contigs, 11, 21, "10")
    loaded = ctx.index_load_mmi(os.path.join(tmp, "ref.mmi"))
    assert loaded.seq_names() == ["chr1", "chr2", "chr3"]
    for a, b in zip(built.export(), loaded.export()):
        assert np.array_equal(a, b)
    o = maplib.sr_opt(min_cnt=0.2, rec_frac=0.1)
    off, lens, buf = flat_reads(reads)
    r1 = ctx.sr_map_batch(built, off, lens, buf, o)
    r2 = ctx.sr_map_batch(loaded, off, lens, buf, o)
    for a, b in zip(r1, r2):
        assert np.array_equal(a, b)
    assert len(r1[1]) > 400
    built.close()
    loaded.close()


def test_sr_map_error_paths_and_capacity_retry(ctx, gd):
    """Argument errors are reported (never a silent different answer) and a too-small output buffer yields
    GD_ERR_CAPACITY with the required sizes, after which the same call succeeds."""
    contigs, reads = maplib.make_dataset(seed=15, n_reads=600)
    idx = ctx.index_build(contigs, 11, 21, "10")
    off, lens, buf = flat_reads(reads)
    o = maplib.sr_opt(min_cnt=0.2, rec_frac=0.1)
    want = ctx.sr_map_batch(idx, off, lens, buf, o)
    got = ctx.sr_map_batch(idx, off, lens, buf, o, cand_cap=8, cigar_cap=8)      # wrapper retries with the reported sizes
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    bad = maplib.sr_opt(af_max_loc=0)
    with pytest.raises(gd.GdietError):
        ctx.sr_map_batch(idx, off, lens, buf, bad)
    short = lens.copy()
    short[5] = 1                                                                  # shorter than the pattern
    with pytest.raises(gd.GdietError):
        ctx.sr_map_batch(idx, off, short, buf, o)
    with pytest.raises(gd.GdietError):
        ctx.index_load_mmi("/nonexistent/file.mmi")
    again = ctx.sr_map_batch(idx, off, lens, buf, o)                              # the context is still usable
    for a, b in zip(want, again):
        assert np.array_equal(a, b)
    idx.close()


# ---- the post-DP stage on the device (gd_sr_map_sam_batch) ----------------------------------------------------------------
@pytest.mark.parametrize("seed,okw,pkw,ragged", [
    (1, {}, {}, False),
    (2, dict(min_cnt=0.2, rec_frac=0.1), dict(no_print_2nd=0, best_n=5), False),
    (7, dict(min_cnt=0.2, bw_min=500, bw_max=1500, af_max_loc=2), dict(softclip=1), True),
    (8, dict(min_cnt=0.2, mid_occ=2, max_max_occ=3, occ_dist=40), dict(sam_hit_only=1, no_print_2nd=0), False),
    (13, dict(min_cnt=0.2, rec_frac=0.1, bw_frac=0.25, bw_min=20, bw_max=60), {}, True),
])
def test_device_sam_stage_equals_host_stage(ctx, seed, okw, pkw, ragged):
    """reads -> SAM text entirely on the device must be byte-identical to gd_sr_map_batch + the host stage gd_sr_sam_batch (which
    is pinned against the reference program's SAM): multi-candidate reads, secondaries printed, soft clips, N, ragged
    lengths, lower case, with and without qualities, several slices and both lanes"""
    import gdiet_b200 as gd
    rng = np.random.default_rng(seed)
    contigs, reads = maplib.make_dataset(seed=seed, n_reads=70000 if seed == 1 else 4000)
    rl = list(reads)
    if ragged:
        rl = [r[:int(rng.integers(30, 151))].copy() for r in rl]
        rl = [np.frombuffer(bytes(r).lower(), np.uint8).copy() if i % 5 == 0 else r for i, r in enumerate(rl)]
    o = maplib.sr_opt(**okw)
    idx = ctx.index_build(contigs, 11, 21, "10")
    off, lens, buf = flat_ragged(rl)
    qual = (33 + (np.arange(len(buf)) % 41)).astype(np.uint8)
    names = ["q%d/x" % i for i in range(len(rl))]
    seq_names = ["chr%d" % (i + 1) for i in range(len(contigs))]
    post = gd.sr_post_options(n_threads=4, **pkw)
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    for q in (qual, None):
        want = gd.sr_sam_batch(names, off, lens, buf, q, coff, cand, cig, seq_names, contigs, post)
        got = ctx.sr_map_sam_batch(idx, names, off, lens, buf, q, o, post, seq_names)
        assert len(got) == len(want) and got == want
    # pieces of the previous call stay valid during the next one (two generations)
    p1 = ctx.sr_map_sam_batch(idx, names, off, lens, buf, qual, o, post, seq_names, join=False)
    first = b"".join(ctypes.string_at(a, l) for a, l in p1)
    ctx.sr_map_sam_batch(idx, names[:100], off[:100], lens[:100], buf, qual, o, post, seq_names, join=False)
    assert b"".join(ctypes.string_at(a, l) for a, l in p1) == first
    idx.close()


def test_pattern_of_length_one_is_refused(ctx):
    """W = 1: mm_sketch2 walks one shift more than the pattern has (sketch.c:2143-2225); not pinned against the reference, so
    the mapping stage and the read-sketch entry refuse it instead of answering differently"""
    import gdiet_b200 as gd
    contigs, reads = maplib.make_dataset(seed=1, n_reads=8)
    idx = ctx.index_build(contigs, 11, 21, "1")
    off, lens, buf = flat_reads(reads)
    o = maplib.sr_opt(Z="1")
    with pytest.raises(gd.GdietError):
        ctx.sr_map_batch(idx, off, lens, buf, o)
    with pytest.raises(gd.GdietError):
        ctx.sketch_reads_batch(off, lens, buf, 11, 21, "1", 0.1, 800)
    idx.close()


def test_device_sam_stage_unmappable_and_empty(ctx):
    import gdiet_b200 as gd
    contigs, _ = maplib.make_dataset(seed=1, n_reads=1)
    idx = ctx.index_build(contigs, 11, 21, "10")
    o, post = maplib.sr_opt(), gd.sr_post_options(n_threads=2)
    junk = np.frombuffer(b"ACGT" * 40, np.uint8)[None, :150].repeat(5, 0).copy()
    junk[2] = np.frombuffer(b"N" * 150, np.uint8)
    off, lens, buf = flat_reads(junk)
    names = ["j%d" % i for i in range(5)]
    got = ctx.sr_map_sam_batch(idx, names, off, lens, buf, None, o, post, ["chr1", "chr2", "chr3"])
    coff, cand, cig = ctx.sr_map_batch(idx, off, lens, buf, o)
    assert got == gd.sr_sam_batch(names, off, lens, buf, None, coff, cand, cig, ["chr1", "chr2", "chr3"], contigs, post)
    assert got.count(b"\n") == 5 and b"\t4\t*\t0\t0\t*" in got
    assert ctx.sr_map_sam_batch(idx, [], np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(1, np.uint8), None, o, post, ["chr1", "chr2", "chr3"]) == b""
    idx.close()
