"""Generates tests/golden/map_sr.npz from the UNMODIFIED reference program (oracle/_ref/GDiet_avx_sr, the GDiet_avx
build with the call tracer of oracle/ref_trace.c): per read, the candidate windows handed to mm_update_extra
(GDiet-ShortReads/map.c:932-954) with the score and CIGAR of exact_match_sse / ksw_extd2_avx512.
Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import maplib  # noqa: E402

SEED, N_READS, MIN_CNT, REC = 31, 2000, 0.2, 0.1
contigs, reads = maplib.make_dataset(seed=SEED, n_reads=N_READS)
o = maplib.sr_opt(min_cnt=MIN_CNT, rec_frac=REC)
sam, tr = maplib.run_reference(contigs, reads, maplib.ref_cmdline(o, extra=["-r", "0.05,150,200", "-n", "%g,%g" % (MIN_CNT, REC)]))
fields = {f: [] for f in ("rid", "rs", "re", "qs", "qe", "rev", "exact", "score", "n_cigar")}
cand_off, cigs = [0], []
for t in tr:
    for c in t["cands"]:
        for f in ("rid", "rs", "re", "qs", "qe", "rev", "exact", "score"):
            fields[f].append(c[f])
        fields["n_cigar"].append(len(c["cigar"]))
        cigs.append(c["cigar"])
    cand_off.append(cand_off[-1] + len(t["cands"]))
np.savez_compressed(os.path.join(HERE, "map_sr.npz"), seed=SEED, n_reads=N_READS, min_cnt=MIN_CNT, rec_frac=REC,
                    cand_off=np.array(cand_off, np.int64), cigar=np.concatenate(cigs).astype(np.uint32),
                    **{f: np.array(v, np.int32) for f, v in fields.items()})
sam_lines = [l for l in sam.splitlines() if not l.startswith("@PG")]
with open(os.path.join(HERE, "map_sr.sam"), "w") as f:
    f.write("\n".join(sam_lines) + "\n")
print("candidates", cand_off[-1], "sam lines", len(sam_lines))
