# sketch tile body v3: parity tests, kernel timings, ncu captures of the contig and the read kernel (round 2)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sketch.py -x -q > gpurun_out/sk3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sk3_pytest.log
tail -3 gpurun_out/sk3_pytest.log
GDIET_SK_V=3 python tools/sketch_bench.py 400 2000000 > gpurun_out/sk3_bench_v3.jsonl 2> gpurun_out/sk3_bench_v3.err
cat gpurun_out/sk3_bench_v3.jsonl
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gd_sketch_tile3 -c 1 -f -o gpurun_out/sk3_ncu python tools/sketch_bench.py 200 1000 > gpurun_out/sk3_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
