#!/bin/bash
# Round-2 session k: the whole GPU suite, smoke, then the driver's own commands (reference arm, our arm) with the final defaults.
TAG=${1:-r2k}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
python -m pytest tests -m gpu -q --timeout 1500 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; tail -4 gpurun_out/bench_reference_$TAG.err; cut -c1-600 gpurun_out/bench_reference_$TAG.json
( time python bench.py --gpus 1 --steps 10 --warmup 3 ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -4 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("C2 value %.1f ms/step %.1f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"], "cpu", d["cpu_baseline"])
for k, v in d["roofline_stages"].items(): print("  ", k, "share %.3f frac %.3f dram_frac %s" % (v["share_of_step"], v["frac"], v.get("dram_frac")))
for wl, v in d["workloads"].items(): print(wl, {k: (round(x, 2) if isinstance(x, float) else x) for k, x in v.items() if k not in ("stages", "clocks", "note")}, v.get("clocks", {}).get("sm_mhz"))
PY
