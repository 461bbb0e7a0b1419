mkdir -p gpurun_out
set -x
python bench.py --steps 2 --warmup 3 --pairs 131072 --flag 0x0 --group 4 --no-cpu > gpurun_out/r7_plain0.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gd_ksw_dp -s 4 -c 1 -o gpurun_out/r7_prof_exact python bench.py --steps 2 --warmup 3 --pairs 131072 --flag 0x0 --group 4 --no-cpu > gpurun_out/r7_ncu0.log 2>&1
python bench.py --steps 2 --warmup 3 --pairs 131072 --flag 0x8 --group 4 --no-cpu > gpurun_out/r7_plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gd_ksw_dp -s 4 -c 1 -o gpurun_out/r7_prof_approx python bench.py --steps 2 --warmup 3 --pairs 131072 --flag 0x8 --group 4 --no-cpu > gpurun_out/r7_ncu8.log 2>&1
ls -la gpurun_out
