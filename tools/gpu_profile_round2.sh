# usage: bash tools/gpu_profile_round2.sh TAG -- end-of-round evidence (round 2): bench lines (default, live flag, reference arm), ncu
# launch list of the bench command, ncu full capture of the DP kernel, launch list + full capture of the device SAM stage
mkdir -p gpurun_out
T=$1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sketch --no-map-strong"
$CMD > gpurun_out/${T}_bench_plain.json 2> gpurun_out/${T}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_bench.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gd_ksw_dp -s 4 -c 1 -o gpurun_out/${T}_dp_full $CMD > gpurun_out/${T}_ncu_full.log 2>&1
python tools/map_sam_kernel_times.py > gpurun_out/${T}_map_sam_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_map_sam.csv python tools/map_sam_kernel_times.py > gpurun_out/${T}_ncu_map_sam.log 2>&1
python tools/map_sam_kernel_times.py > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gd_sam_write -s 4 -c 1 -o gpurun_out/${T}_sam_write_full python tools/map_sam_kernel_times.py > gpurun_out/${T}_ncu_sam.log 2>&1
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
python bench.py --flag 0x8 --no-cpu --no-sketch --no-map-strong > gpurun_out/${T}_bench_flag8.json 2>> gpurun_out/${T}_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench_default.err
ls -la gpurun_out | tail -14
