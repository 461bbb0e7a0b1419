# sketch tile body v3 against v2: parity tests, kernel timings, one ncu capture (round 2)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sketch.py -x -q > gpurun_out/sk3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sk3_pytest.log
tail -3 gpurun_out/sk3_pytest.log
for V in 3 2; do
  GDIET_SK_V=$V python tools/sketch_bench.py 400 2000000 > gpurun_out/sk3_bench_v$V.jsonl 2> gpurun_out/sk3_bench_v$V.err
  cat gpurun_out/sk3_bench_v$V.jsonl
done
GDIET_SK_V=3 GDIET_SK_THREADS=128 python tools/sketch_bench.py 400 1000 > gpurun_out/sk3_bench_v3_t128.jsonl 2>&1
head -1 gpurun_out/sk3_bench_v3_t128.jsonl
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gd_sketch_tile3 -c 1 -f -o gpurun_out/sk3_ncu python tools/sketch_bench.py 200 1000 > gpurun_out/sk3_ncu.log 2>&1
ls -la gpurun_out/sk3_ncu.ncu-rep
