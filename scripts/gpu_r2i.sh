#!/bin/bash
# Round-2 session i: sweep of the generation-order block (BTRC_GEN_BLOCK=<w>x<h>[s]) and of the state pool on C2 / C3.
TAG=${1:-r2i}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
run() {  # name, env assignments..., -- bench args
  NAME=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads $ARGS > gpurun_out/sw_${NAME}_$TAG.json 2> gpurun_out/sw_${NAME}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sw_${NAME}_$TAG.json"))
    print("$NAME", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
except Exception as e:
    print("$NAME failed", e)
PY
}
ARGS="--workload c2_disney_1080p_256spp --spp 64"
for B in 8x8 16x16 32x16 64x32 128x64 256x128 8x8s 32x16s 128x64s; do run c2_$B BTRC_GEN_BLOCK=$B; done
ARGS="--workload c2_disney_1080p_256spp --spp 64 --state-count 4194304"; run c2_pool4M BTRC_GEN_BLOCK=32x16
ARGS="--workload c2_disney_1080p_256spp --spp 64 --state-count 16777216"; run c2_pool16M BTRC_GEN_BLOCK=32x16
ARGS="--workload c2_disney_1080p_256spp --spp 64 --state-count 2097152"; run c2_pool2M BTRC_GEN_BLOCK=32x16
ARGS="--workload c3_medium_1080p_128spp --spp 32"
for B in 8x8 32x16 128x64 32x16s; do run c3_$B BTRC_GEN_BLOCK=$B; done
