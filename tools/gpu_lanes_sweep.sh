# usage: bash tools/gpu_lanes_sweep.sh -- the mapping stage (10 M reads vs 3.1 Gbp, one GPU) against the number of lanes
mkdir -p gpurun_out
for l in 2 3 4 6; do
  GDIET_MAP_LANES=$l python tools/map_strong_bench.py 10000000 3.1 > gpurun_out/r2_lanes_$l.json 2> gpurun_out/r2_lanes_$l.err || tail -3 gpurun_out/r2_lanes_$l.err
done
python - <<'PY'
import json
for l in (2, 3, 4, 6):
    try:
        d = json.loads(open("gpurun_out/r2_lanes_%d.json" % l).read().strip().splitlines()[-1])
        print(l, "map_call %.1f M reads/s" % (d["map_call"]["reads_per_s"] / 1e6), "e2e %.1f" % (d["e2e"]["reads_per_s"] / 1e6), "host-stage %.1f" % (d["e2e_host_stage"]["reads_per_s"] / 1e6),
              d["e2e"].get("identical_to_host_stage_on_ranks"), d.get("sam_sha256", "")[:12])
    except Exception as e:
        print(l, "failed", e)
PY
