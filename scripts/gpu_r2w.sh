#!/bin/bash
# Round-2 session w: generate kernel with 32-bit index arithmetic and per-bit PCG skip-ahead -- parity tests that exercise every index map, then C2.
TAG=${1:-r2w}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x -k "generate or partition or ragged or c5 or 2d_tiles or parity or halton or tile or c1 or converged" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_c2_$TAG.json"))
print("C2 value %.1f" % d["value"], {k: (round(v["ms_per_launch"], 3), round(v["frac"], 3)) for k, v in d["roofline_stages"].items()})
PY
