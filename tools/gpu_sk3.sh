# sketch tile body v3: parity tests, kernel timings with the A/B switches (packed read tiles, early ticket) (round 2)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sketch.py tests/test_gpu_map.py -x -q > gpurun_out/sk3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sk3_pytest.log
tail -3 gpurun_out/sk3_pytest.log
python tools/sketch_bench.py 400 2000000 > gpurun_out/sk3_bench_default.jsonl 2> gpurun_out/sk3_bench_default.err
cat gpurun_out/sk3_bench_default.jsonl
GDIET_SK_PACK=0 GDIET_SK_EARLY=0 python tools/sketch_bench.py 400 2000000 > gpurun_out/sk3_bench_p0e0.jsonl 2>&1
cat gpurun_out/sk3_bench_p0e0.jsonl
GDIET_SK_PACK=1 GDIET_SK_EARLY=0 python tools/sketch_bench.py 25 2000000 > gpurun_out/sk3_bench_p1e0.jsonl 2>&1
tail -1 gpurun_out/sk3_bench_p1e0.jsonl
