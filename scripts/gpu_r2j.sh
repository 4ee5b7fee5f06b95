#!/bin/bash
# Round-2 session j: pool size at full spp with the coherent order, if-if on C2, bigger pool on C4.
TAG=${1:-r2j}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
run() {
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
ARGS="--workload c2_disney_1080p_256spp --state-count 8388608"; run c2_8M X=1
ARGS="--workload c2_disney_1080p_256spp --state-count 16777216"; run c2_16M X=1
ARGS="--workload c2_disney_1080p_256spp --state-count 33554432"; run c2_32M X=1
ARGS="--workload c2_disney_1080p_256spp --state-count 16777216"; run c2_16M_ifif BTRC_TRAVERSE=ifif
ARGS="--workload c2_disney_1080p_256spp --state-count 16777216"; run c2_16M_sort BTRC_SORT_RAYS=2
ARGS="--workload c3_medium_1080p_128spp --state-count 16777216"; run c3_16M X=1
ARGS="--workload c4_instanced_10m_1080p_128spp --state-count 33554432"; run c4_32M X=1
ARGS="--workload c4_instanced_10m_1080p_128spp --state-count 16777216"; run c4_16M_ww BTRC_TRAVERSE=ww
ARGS="--workload c1_cornell_512_16spp --state-count 4194304 --steps 100"; run c1_4M X=1
ARGS="--workload c1_cornell_512_16spp --state-count 1048576 --steps 100"; run c1_1M X=1
