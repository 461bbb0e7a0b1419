#!/usr/bin/env python
"""Host stage alone (gd_sr_sam_batch: mm_update_extra ... SAM records), no GPU needed: synthetic 150 bp reads with one
150M candidate each against a 200 Mbp genome; prints seconds and microseconds per read and thread.
    python tools/sam_stage_bench.py [n_reads] [threads]        (GDIET_SAM_HUGEPAGES=0 to compare page sizes)"""
import sys, time, ctypes as C, numpy as np
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import gdiet_b200 as gd
n=int(sys.argv[1]) if len(sys.argv)>1 else 1000000
threads=int(sys.argv[2]) if len(sys.argv)>2 else 1
G=200_000_000
rng=np.random.default_rng(1)
genome=np.frombuffer(b"ACGT",np.uint8)[rng.integers(0,4,G,dtype=np.uint8)]
pos=rng.integers(0,G-200,n)
reads=genome[pos[:,None]+np.arange(150)[None,:]]
sub=rng.random((n,150))<0.01
reads=np.where(sub, np.frombuffer(b"ACGT",np.uint8)[rng.integers(0,4,(n,150),dtype=np.uint8)], reads).astype(np.uint8)
rev=rng.random(n)<0.5
comp=np.zeros(256,np.uint8); 
for a,b in zip(b"ACGT",b"TGCA"): comp[a]=b
reads[rev]=comp[reads[rev][:,::-1]]
buf=np.ascontiguousarray(reads.reshape(-1)); off=np.arange(n,dtype=np.int64)*150; lens=np.full(n,150,np.int32)
qual=np.full(n*150,ord('I'),np.uint8)
cand=np.zeros(n,gd.SR_CAND_DTYPE)
cand["rid"]=0; cand["rs"]=pos; cand["re"]=pos+150; cand["qs"]=0; cand["qe"]=150; cand["rev"]=rev; cand["score"]=280; cand["n_cigar"]=1
cand["cigar_off"]=np.arange(n)
cig=np.full(n,150<<4,np.uint32)
coff=np.arange(n+1,dtype=np.int64)
sys.path.insert(0,os.path.join(ROOT,'tools'))
import map_strong_bench as msb
nb,ptr=msb.fixed_names(n)
names=(C.c_char_p*n).from_buffer(ptr)
post=gd.sr_post_options(n_threads=threads)
ref=(np.zeros(1,np.int64),np.array([G],np.int32),genome)
for it in range(3):
    t0=time.perf_counter()
    p=gd.sr_sam_batch(names,off,lens,buf,qual,coff,cand,cig,["chr1"],None,post,parts=True,ref=ref)
    dt=time.perf_counter()-t0
    print("threads",threads,"reads",n,"s",round(dt,3),"us/read/thread",round(dt*threads/n*1e6,3),"bytes",p.n)
    p.free()
