#!/bin/bash
# Round-2 GPU session c: whole parity suite (no -x), the traversal-heavy tests again on the two-level BVH layout, and A/B bench
# lines flat vs instanced layout on C4 and C2.
TAG=${1:-r2c}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
python -m pytest tests -m gpu -q --timeout 1500 -s 2>&1 | grep -v "^\s*$" | tail -120 > gpurun_out/pytest_gpu_$TAG.log; tail -5 gpurun_out/pytest_gpu_$TAG.log
BTRC_BVH_LAYOUT=instanced python -m pytest tests/test_gpu_parity.py tests/test_gpu_instanced.py tests/test_gpu_materials_media.py tests/test_gpu_volumes.py -m gpu -q --timeout 1500 2>&1 | tail -15 > gpurun_out/pytest_gpu_instanced_$TAG.log; tail -5 gpurun_out/pytest_gpu_instanced_$TAG.log
for L in flat instanced; do
  BTRC_BVH_LAYOUT=$L python bench.py --workload c4_instanced_10m_1080p_128spp --steps 2 --warmup 3 --state-count 16777216 --no-cpu-baseline > gpurun_out/bench_c4_${L}_$TAG.json 2> gpurun_out/bench_c4_${L}_$TAG.err
  BTRC_BVH_LAYOUT=$L python bench.py --workload c2_disney_1080p_256spp --spp 64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_${L}_$TAG.json 2> gpurun_out/bench_c2_${L}_$TAG.err
done
python - <<PY
import json
for wl in ("c4", "c2"):
    for L in ("flat", "instanced"):
        try:
            d = json.load(open("gpurun_out/bench_%s_%s_$TAG.json" % (wl, L)))
            print(wl, L, "value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
        except Exception as e:
            print(wl, L, "failed", e)
PY
