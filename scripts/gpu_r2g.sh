#!/bin/bash
# Round-2 session g: block-coherent generation order A/B on all single-GPU workloads + the full parity suite with it on.
TAG=${1:-r2g}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
python -m pytest tests -m gpu -q --timeout 1500 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
for O in 0 1; do
  BTRC_GEN_ORDER=$O python bench.py --workload c2_disney_1080p_256spp --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_c2_order${O}_$TAG.json 2> gpurun_out/bench_c2_order${O}_$TAG.err
  BTRC_GEN_ORDER=$O python bench.py --workload c1_cornell_512_16spp --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1_order${O}_$TAG.json 2> gpurun_out/bench_c1_order${O}_$TAG.err
  BTRC_GEN_ORDER=$O python bench.py --workload c3_medium_1080p_128spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_order${O}_$TAG.json 2> gpurun_out/bench_c3_order${O}_$TAG.err
  BTRC_GEN_ORDER=$O python bench.py --workload c4_instanced_10m_1080p_128spp --steps 2 --warmup 3 --state-count 16777216 --no-cpu-baseline > gpurun_out/bench_c4_order${O}_$TAG.json 2> gpurun_out/bench_c4_order${O}_$TAG.err
done
python - <<PY
import json
for wl in ("c2", "c1", "c3", "c4"):
    for O in (0, 1):
        try:
            d = json.load(open("gpurun_out/bench_%s_order%d_$TAG.json" % (wl, O)))
            print(wl, "order", O, "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
        except Exception as e:
            print(wl, O, "failed", e)
PY
