"""Launch-list workload for ncu: 1 M reads x 150 bp against 50 Mbp through gd_sr_map_sam_batch (reads in, SAM text out, one lane), three
passes; prints wall seconds per pass.  Used by tools/gpu_profile_round2.sh."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, gdiet_b200 as gd
from gdiet_b200 import synth
n = 1_000_000
genome = synth.random_genome(50_000_000, seed=1)
reads = synth.sample_reads(genome, n, 150, seed=2)
off = np.arange(n, dtype=np.int64) * 150
lens = np.full(n, 150, np.int32)
buf = np.ascontiguousarray(reads.reshape(-1))
qual = np.full(n * 150, ord("I"), np.uint8)
names = gd._cstr_array(["r%d" % i for i in range(n)])
ctx = gd.Context(0)
idx = ctx.index_build([genome], 11, 21, "10")
opt, post = gd.sr_options(), gd.sr_post_options()
ctx.set_option("map_lanes", 1)
for it in range(3):
    t0 = time.perf_counter()
    p = ctx.sr_map_sam_batch(idx, names, off, lens, buf, qual, opt, post, ["chr1"], join=False)
    print(json.dumps({"map_sam_s": round(time.perf_counter() - t0, 4), "sam_bytes": sum(l for _, l in p)}))
