#!/bin/bash
# Round-2 session t: full verification with the final defaults (scripts/gpu_r2k.sh), then shade launch-bound variants on C2 64 spp.
TAG=${1:-r2t}
bash scripts/gpu_r2k.sh $TAG
for V in default s128x5 s192x3 default; do
  L=btrc_b200/libbtrc_b200.so; [ $V != default ] && L=btrc_b200/variants/libbtrc_$V.so
  [ -f $L ] || continue
  BTRC_B200_LIB=$PWD/$L python bench.py --workload c2_disney_1080p_256spp --spp 64 --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_c2_${V}_$TAG.json 2> gpurun_out/bench_c2_${V}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c2_${V}_$TAG.json"))
    print("$V", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
except Exception as e:
    print("$V failed", e)
PY
done
