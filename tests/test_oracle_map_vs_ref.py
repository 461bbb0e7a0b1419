"""Pins oracle/gd_oracle_map.c (index lookup, shift selection, seed filters, hit sort, voting, window arithmetic:
SURVEY.md 8 rows F1/F2) against the UNMODIFIED reference program: oracle/_ref/GDiet_avx_sr is the GDiet_avx build
with a call tracer (oracle/ref_trace.c) in front of exact_match_sse / ksw_extd2_avx512 / mm_update_extra, so every
candidate window, its query/target bytes, score and CIGAR of every read is compared."""
import os

import numpy as np
import pytest

import maplib
from oraclelib import cpu_has_avx512

pytestmark = pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()),
                                reason="needs oracle/_ref/GDiet_avx_sr (built where /root/reference exists) and AVX-512")

CASES = [
    # seed, Z, read_len, extra command-line flags, option overrides
    (1, "10", 150, ["-r", "0.05,150,200"], {}),                                               # BASELINE config 1 flags
    (2, "10", 150, ["-r", "0.05,150,200", "-n", "0.2,0.1"], dict(min_cnt=0.2, rec_frac=0.1)),  # real votes + recovery threshold
    (3, "110", 150, ["-r", "0.05,150,200", "-n", "0.3"], dict(min_cnt=0.3)),
    (4, "10", 400, ["-n", "0.2"], dict(min_cnt=0.2, bw_min=500, bw_max=1500)),                 # qlen > 300: seed-bounded windows
    (5, "10", 100, ["-r", "0.1,20,50", "-n", "0.1"], dict(min_cnt=0.1, bw_frac=0.1, bw_min=20, bw_max=50)),  # narrow band
    (7, "10", 300, ["-n", "0.2", "--AF_max_loc", "2"], dict(min_cnt=0.2, bw_min=500, bw_max=1500, af_max_loc=2)),
]


@pytest.fixture(scope="module")
def M():
    return maplib.MapOracle()


@pytest.mark.parametrize("seed,Z,read_len,extra,okw", CASES)
def test_map_oracle_matches_reference_trace(M, seed, Z, read_len, extra, okw):
    contigs, reads = maplib.make_dataset(seed=seed, read_len=read_len, n_reads=1200)
    o = maplib.sr_opt(Z=Z, qlen=read_len, **okw)
    _, tr = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=extra))
    assert len(tr) == len(reads)
    mi = M.index_build(contigs, 11, 21, Z)
    n_cand = 0
    for i, (r, t) in enumerate(zip(reads, tr)):
        assert bytes(r) == t["seq"]
        c, cig, _ = M.map_read(mi, r, o)
        maplib.cands_equal_trace(c, cig, t["cands"], "seed %d read %d" % (seed, i))
        n_cand += len(c)
    M.lib.gdo_index_destroy(mi)
    assert n_cand > len(reads) // 2


def test_map_oracle_mixed_read_lengths_per_read_band(M):
    """-r 0.25,20,60 on reads of 60..300 bases: the clamp of map.c:624-631 is not saturated, so the band and the vote
    distance differ from read to read inside one run"""
    rng = np.random.default_rng(5)
    contigs, base = maplib.make_dataset(seed=12, read_len=300, n_reads=900)
    reads = [r[:int(rng.integers(60, 301))].copy() for r in base]
    o = maplib.sr_opt(min_cnt=0.2, rec_frac=0.1, bw_frac=0.25, bw_min=20, bw_max=60)
    _, tr = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=["-r", "0.25,20,60", "-n", "0.2,0.1"]))
    mi = M.index_build(contigs, 11, 21, "10")
    bands, n_cand = set(), 0
    for i, (r, t) in enumerate(zip(reads, tr)):
        c, cig, _ = M.map_read(mi, r, o)
        maplib.cands_equal_trace(c, cig, t["cands"], "read %d (len %d)" % (i, len(r)))
        bands.update(int(x["w"]) for x in t["cands"] if not x["exact"])
        n_cand += len(c)
    M.lib.gdo_index_destroy(mi)
    assert n_cand > 400 and len(bands) > 20 and min(bands) == 20 and max(bands) == 60


def test_reference_sam_is_deterministic_across_threads():
    """the SAM (minus @PG) does not depend on -t; the golden SAM fixtures rely on it."""
    contigs, reads = maplib.make_dataset(seed=1, n_reads=400)
    o = maplib.sr_opt()
    s1, _ = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=["-r", "0.05,150,200"]), threads=1, trace=False)
    s4, _ = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=["-r", "0.05,150,200"]), threads=4, trace=False)
    strip = lambda s: [l for l in s.splitlines() if not l.startswith("@PG")]
    assert strip(s1) == strip(s4)
