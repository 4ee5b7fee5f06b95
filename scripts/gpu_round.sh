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
# full captures: launches #36..#39 of each heavy kernel (steady state of a warm-up render); BTRC_ITER_LOG gives the ray count of every launch
export BTRC_ITER_LOG=1
for K in trace_kernel shade_kernelILi2E shadow_kernel; do
  N=${K%%_kernel*}
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s 36 -c 4 -f -o gpurun_out/prof_${N}_$TAG \
      $PROF > gpurun_out/ncu_${N}_$TAG.log 2>&1
  grep "btrc iter" gpurun_out/ncu_${N}_$TAG.log > gpurun_out/iters_${N}_$TAG.log
  python scripts/ncu_summary.py traffic gpurun_out/prof_${N}_$TAG.ncu-rep gpurun_out/iters_${N}_$TAG.log 36 $N gpurun_out/traffic_$WL.json > /dev/null
done
cat gpurun_out/traffic_$WL.json | head -50
ls -la gpurun_out
