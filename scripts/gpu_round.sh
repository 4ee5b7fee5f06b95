#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list + full captures of the two heavy kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 3 -f -o gpurun_out/prof_trace_$TAG \
    python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_trace_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shade_kernel -s 2 -c 2 -f -o gpurun_out/prof_shade_$TAG \
    python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_shade_$TAG.log 2>&1
ls -la gpurun_out
