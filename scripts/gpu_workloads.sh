#!/bin/bash
# Bench lines of the non-default BASELINE configs on one GPU (parity-test scenes; the default bench line is C2).  Usage: bash scripts/gpu_workloads.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --workload c1_cornell_512_16spp --steps 10 --warmup 3 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err
python bench.py --workload c3_medium_1080p_128spp --steps 3 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
python bench.py --workload c4_instanced_10m_1080p_128spp --steps 3 --warmup 3 --state-count 16777216 --no-cpu-baseline > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
for C in c1 c3 c4; do
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${C}_$TAG.json"))
    print("$C", d["config"]["workload"], "value %.1f Msamples/s  mrays %.1f  ms/step %.2f  e2e %.1f" % (d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"]), "cpu", d["cpu_baseline"])
    for k, v in d["roofline_stages"].items():
        print("   %-12s ms/launch %.4f share %.3f  GB/s %.0f frac %.3f" % (k, v["ms_per_launch"], v["share_of_step"], v["achieved_gbs"], v["frac"]))
except Exception as e:
    print("$C failed", e)
PY
tail -3 gpurun_out/bench_${C}_$TAG.err
done
