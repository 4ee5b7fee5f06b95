#!/bin/bash
# Round-2 GPU session d: full parity suite with the complete log, C3 bench with the batched tracking loops (A/B against BTRC_TRACK_BATCH=1).
TAG=${1:-r2d}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
python -m pytest tests -m gpu -q --timeout 1500 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
python bench.py --workload c3_medium_1080p_128spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_batch4_$TAG.json 2> gpurun_out/bench_c3_batch4_$TAG.err
cp btrc_b200/libbtrc_b200.so /tmp/lib_batch4.so
BTRC_NVCC_EXTRA="-DBTRC_TRACK_BATCH=1" python -m btrc_b200.build --force > /dev/null 2>&1
python bench.py --workload c3_medium_1080p_128spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_batch1_$TAG.json 2> gpurun_out/bench_c3_batch1_$TAG.err
BTRC_NVCC_EXTRA="-DBTRC_TRACK_BATCH=8" python -m btrc_b200.build --force > /dev/null 2>&1
python bench.py --workload c3_medium_1080p_128spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_batch8_$TAG.json 2> gpurun_out/bench_c3_batch8_$TAG.err
python - <<PY
import json
for b in ("batch1", "batch4", "batch8"):
    try:
        d = json.load(open("gpurun_out/bench_c3_%s_$TAG.json" % b))
        print("c3", b, "value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(b, "failed", e)
PY
