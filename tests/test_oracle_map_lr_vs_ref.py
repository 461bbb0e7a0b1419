"""Pins the long-read part of oracle/gd_oracle_map.c (two voting rounds, density / score filters, candidate chaining,
windows: GDiet-LongReads/map.c:1052-1805) against the call trace of the UNMODIFIED long-read reference program
(oracle/_ref/GDiet_avx_lr + oracle/ref_trace.c): every DP call of every read (lengths, score, CIGAR) and the candidate
window handed to mm_update_extra."""
import pytest

import gdiet_b200 as gd
import maplib
from oraclelib import cpu_has_avx512

pytestmark = pytest.mark.skipif(not (maplib.have_ref_program() and cpu_has_avx512()),
                                reason="needs oracle/_ref/GDiet_avx_lr (built where /root/reference exists) and AVX-512")

ONT_FLAGS = ["--vt_dis=1000", "--vt_nb_loc=3", "--vt_df1=0.007", "--vt_df2=0.007", "--max_min_gap=4000", "--vt_f=0.04", "--vt_cov",
             "0.3", "--sort=merge", "--frag=no"]   # the flag set of the reference's README for ONT reads (SURVEY.md finding 6)
ONT_OPT = dict(vt_dis=1000, vt_df1=0.007, vt_df2=0.007, vt_f=0.04, vt_cov=0.3)

LR_CASES = [
    # seed, preset, k, w, bw, read_len, sub, indel, n_reads, extra flags, option overrides
    (1, "map-hifi", 19, 19, 1000, 15000, 0.005, 0.005, 24, [], {}),
    (3, "map-hifi", 19, 19, 400, 6000, 0.005, 0.005, 40, ["--vt_nb_loc=2"], dict(vt_nb_loc=2)),
    (2, "map-ont", 15, 10, 1300, 12000, 0.03, 0.05, 24, ONT_FLAGS, ONT_OPT),
]


def lr_setup(M, seed, preset, k, w, bw, read_len, sub, indel, n_reads, extra, okw):
    contigs, reads = maplib.make_long_dataset(seed=seed, read_len=read_len, sub=sub, indel=indel, n_reads=n_reads)
    flags = ["-ax", preset, "-Z", "10", "-W", "2", "-k", str(k), "-w", str(w), "-r", str(bw)] + list(extra)
    mi = M.index_build(contigs, w, k, "10")
    lo, hi = (50, 500) if preset == "map-hifi" else (10, 1000000)   # min_mid_occ / max_mid_occ, LR/options.c:17-18,108
    mid = min(max(M.lib.gdo_index_cal_max_occ(mi, 2e-4), lo), hi)
    return contigs, reads, flags, mi, gd.lr_options(preset, bw=bw, mid_occ=mid, **okw)


@pytest.mark.parametrize("case", LR_CASES)
def test_lr_map_oracle_matches_reference_trace(case):
    M = maplib.MapOracle()
    contigs, reads, flags, mi, o = lr_setup(M, *case)
    _, tr = maplib.run_reference(contigs, reads, flags, program=maplib.REF_LR)
    assert len(tr) == len(reads)
    n_cand = 0
    for i, (r, t) in enumerate(zip(reads, tr)):
        assert bytes(r) == t["seq"]
        c, cig, _ = M.lr_map_read(mi, r, o)
        maplib.lr_cands_equal_trace(c, cig, t["cands"], "read %d" % i)
        n_cand += len(c)
    M.lib.gdo_index_destroy(mi)
    assert n_cand >= len(reads) // 3
