#!/bin/bash
# Round-2 session o: medium kernel compiled for 2 / 3 / 4 CTAs per SM (prebuilt variants), C3 at a 16 Mi pool.
TAG=${1:-r2o}
mkdir -p gpurun_out
for V in default medium3 medium4 default; do
  L=btrc_b200/libbtrc_b200.so; [ $V != default ] && L=btrc_b200/variants/libbtrc_$V.so
  BTRC_B200_LIB=$PWD/$L python bench.py --workload c3_medium_1080p_128spp --state-count 16777216 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_${V}_$TAG.json 2> gpurun_out/bench_c3_${V}_$TAG.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c3_${V}_$TAG.json"))
    print("$V", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
except Exception as e:
    print("$V failed", e)
PY
done
BTRC_B200_LIB=$PWD/btrc_b200/variants/libbtrc_medium3.so python -m pytest tests/test_gpu_materials_media.py tests/test_gpu_volumes.py -m gpu -q --timeout 900 2>&1 | tail -3
