#!/bin/bash
# Round-2 session y: state pool beyond 32 Mi on C2 and C3 (driver-style command, 3 steps).
TAG=${1:-r2y}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for POOL in 33554432 50331648 67108864; do
  for WL in c2_disney_1080p_256spp c3_medium_1080p_128spp; do
    python bench.py --workload $WL --state-count $POOL --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_${WL%%_*}_pool${POOL}_$TAG.json 2> gpurun_out/bench_${WL%%_*}_pool${POOL}_$TAG.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${WL%%_*}_pool${POOL}_$TAG.json"))
    print("${WL%%_*} pool $POOL value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_total"], 1) for k, v in d["roofline_stages"].items()})
except Exception as e:
    print("${WL%%_*} pool $POOL failed", e, open("gpurun_out/bench_${WL%%_*}_pool${POOL}_$TAG.err").read()[-300:])
PY
  done
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
