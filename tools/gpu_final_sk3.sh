# end-of-round check after the sketch kernel rewrite (v3): GPU suite, smoke, a bounded randomised sketch sweep, the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_sk3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_sk3_pytest.log
tail -3 gpurun_out/r2_sk3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sk3_smoke.log 2>&1; tail -1 gpurun_out/r2_sk3_smoke.log
python tools/sketch_fuzz.py --seconds 10 --seed 20261020 > gpurun_out/r2_sk3_sketch_fuzz.jsonl 2> gpurun_out/r2_sk3_sketch_fuzz.err; tail -1 gpurun_out/r2_sk3_sketch_fuzz.jsonl | cut -c1-200
python bench.py > gpurun_out/r2_sk3_bench_default.json 2> gpurun_out/r2_sk3_bench_default.err; tail -c 600 gpurun_out/r2_sk3_bench_default.json
