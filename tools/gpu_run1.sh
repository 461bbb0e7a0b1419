mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1
for G in 4 8; do for F in 0x8 0x0; do
python bench.py --steps 5 --warmup 3 --pairs 262144 --flag $F --group $G --no-cpu >> gpurun_out/r4_bench_tune.jsonl 2>> gpurun_out/r4_bench_err.log
done; done
python bench.py --steps 5 --warmup 3 > gpurun_out/r4_bench_full.jsonl 2> gpurun_out/r4_bench_full.err
tail -3 gpurun_out/r4_pytest.log; tail -2 gpurun_out/r4_smoke.log
python - <<'PY'
import json
for f in ("gpurun_out/r4_bench_tune.jsonl","gpurun_out/r4_bench_full.jsonl"):
    for l in open(f):
        d=json.loads(l); print(d["config"]["flag"], d["extra"]["ksw_group_lanes"], round(d["value"],1), round(d["e2e"]["value"],1), d["config"]["pairs_per_gpu"])
PY
