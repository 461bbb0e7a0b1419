# usage: bash tools/gpu_profile_final.sh TAG -- end-of-round evidence: bench lines (default + live flag), ncu launch list of the
# bench command, ncu full captures of the DP kernel and of the 256-thread sketch kernel (index build)
mkdir -p gpurun_out
T=$1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sketch"
$CMD > gpurun_out/${T}_bench_plain.json 2> gpurun_out/${T}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gd_ksw_dp -s 4 -c 1 -o gpurun_out/${T}_dp_full $CMD > gpurun_out/${T}_ncu_full.log 2>&1
python tools/sketch_bench.py 400 200000 > gpurun_out/${T}_sketch_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gd_sketch_tile_kernel -s 1 -c 1 -o gpurun_out/${T}_sketch_full python tools/sketch_bench.py 400 200000 > gpurun_out/${T}_ncu_sketch.log 2>&1
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
python bench.py --flag 0x8 --no-cpu > gpurun_out/${T}_bench_flag8.json 2>> gpurun_out/${T}_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench_default.err
ls -la gpurun_out | tail -12
