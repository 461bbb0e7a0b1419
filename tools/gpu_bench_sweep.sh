# usage: bash tools/gpu_bench_sweep.sh TAG [pytest]
mkdir -p gpurun_out
T=$1
if [ "$2" = "pytest" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
fi
for G in 4 8; do for F in 0x8 0x0; do
python bench.py --steps 5 --warmup 3 --pairs 262144 --flag $F --group $G --no-cpu --no-sketch >> gpurun_out/${T}_bench_tune.jsonl 2>> gpurun_out/${T}_bench_err.log
done; done
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_tune.jsonl"):
    d=json.loads(l); print(d["config"]["flag"], d["extra"]["ksw_group_lanes"], round(d["value"],1), round(d["e2e"]["value"],1), d["config"]["pairs_per_gpu"])
PY
tail -5 gpurun_out/${T}_bench_err.log
