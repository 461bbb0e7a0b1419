import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np, gdiet_b200 as gd
from gdiet_b200 import synth
n=1_000_000
genome=synth.random_genome(50_000_000, seed=1); reads=synth.sample_reads(genome,n,150,seed=2)
off=np.arange(n,dtype=np.int64)*150; lens=np.full(n,150,np.int32); buf=np.ascontiguousarray(reads.reshape(-1))
ctx=gd.Context(0); idx=ctx.index_build([genome],11,21,"10"); opt=gd.sr_options()
ctx.set_option("time_kernels",1); ctx.set_option("map_lanes",1)
for it in range(3):
    ctx.stat("sketch_reset")
    t0=time.perf_counter(); r=ctx.sr_map_batch(idx,off,lens,buf,opt,cand_cap=n+1024,cigar_cap=8*n+1024); dt=time.perf_counter()-t0
    print(json.dumps({"map_s":round(dt,4),"sketch_kernel_ms":ctx.stat("sketch_us")/1e3,"dp_kernel_ms":ctx.stat("ksw_dp_us")/1e3,"read_sketch_gbases_s":n*150/(ctx.stat("sketch_us")*1e-6)/1e9}))
