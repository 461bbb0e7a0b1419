#!/bin/bash
# one --set full capture of the read-sketch kernel (gd_sketch_tile_kernel<32>) inside the mapping stage
set -x
timeout 300 python tools/sr_map_bench.py 5 200000 noref > gpurun_out/r26_plain.json 2> gpurun_out/r26_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gd_sketch_tile_kernel -s 2 -c 1 -o gpurun_out/r26_readsketch python tools/sr_map_bench.py 5 200000 noref > gpurun_out/r26_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gd_sr_vote_kernel -s 1 -c 1 -o gpurun_out/r26_vote python tools/sr_map_bench.py 5 200000 noref >> gpurun_out/r26_ncu.log 2>&1
tail -3 gpurun_out/r26_ncu.log
