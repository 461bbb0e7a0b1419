# DP kernel throughput against resident blocks per SM (bench.py --blocks-per-sm): how much a smaller shared-memory footprint would buy
for spec in "0x0 3" "0x0 4" "0x0 5" "0x8 1" "0x8 2" "0x8 3"; do
	set -- $spec
	python bench.py --flag $1 --steps 3 --warmup 3 --no-cpu --no-sketch --no-map-strong --blocks-per-sm $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'flag':'$1','blocks_per_sm':$2,'kernel_gcups':round(d['roofline']['kernel_gcups'],1),'value':round(d['value'],1)}))"
done | tee gpurun_out/r2_occupancy_sweep.jsonl
