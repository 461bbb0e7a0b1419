# sketch tile body v3: parity tests, kernel timings (deferred look-back on / off, v2), ncu captures of the contig and the read kernel (round 2)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sketch.py -x -q > gpurun_out/sk3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sk3_pytest.log
tail -3 gpurun_out/sk3_pytest.log
GDIET_SK_V=3 python tools/sketch_bench.py 400 2000000 > gpurun_out/sk3_bench_v3.jsonl 2> gpurun_out/sk3_bench_v3.err
cat gpurun_out/sk3_bench_v3.jsonl
GDIET_SK_V=3 GDIET_SK_DEFER=0 python tools/sketch_bench.py 400 1000 > gpurun_out/sk3_bench_v3_nodefer.jsonl 2>&1
head -1 gpurun_out/sk3_bench_v3_nodefer.jsonl
GDIET_SK_V=2 python tools/sketch_bench.py 400 1000 > gpurun_out/sk3_bench_v2.jsonl 2>&1
head -1 gpurun_out/sk3_bench_v2.jsonl
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gd_sketch_tile3 -c 1 -f -o gpurun_out/sk3_ncu python tools/sketch_bench.py 200 1000 > gpurun_out/sk3_ncu.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gd_sketch_tile3 --launch-skip 4 -c 1 -f -o gpurun_out/sk3_ncu_reads python tools/sketch_bench.py 25 1000000 > gpurun_out/sk3_ncu_reads.log 2>&1
ls -la gpurun_out/*.ncu-rep
