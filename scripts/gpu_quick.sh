#!/bin/bash
# Quick A/B loop on the GPU box: parity tests + a short bench of the default workload.  Usage: bash scripts/gpu_quick.sh <tag> [spp] [workload]
TAG=${1:-q}
SPP=${2:-32}
WL=${3:-c2_disney_1080p_256spp}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --workload $WL --steps 3 --warmup 3 --spp $SPP --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -5 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.1f Msamples/s  mrays %.1f  ms/step %.2f  e2e %.1f  launches %d" % (d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
for k, v in d["roofline_stages"].items():
    print("  %-12s ms/launch %.4f share %.3f  GB/s %.0f frac %.3f" % (k, v["ms_per_launch"], v["share_of_step"], v["achieved_gbs"], v["frac"]))
PY
