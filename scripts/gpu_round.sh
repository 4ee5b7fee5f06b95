#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list + full captures of the heavy kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [workload]
TAG=${1:-r1}
WL=${2:-c2_disney_1080p_256spp}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --workload $WL --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 4000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
# profiler passes run the same workload at 8 spp (the launch mix of a step is the same; numbers printed under ncu are not bench values)
PROF="python bench.py --workload $WL --steps 1 --warmup 3 --spp 8 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv \
    $PROF > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in trace shade shadow; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 6 -c 2 -f -o gpurun_out/prof_${K}_$TAG \
      $PROF > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out
