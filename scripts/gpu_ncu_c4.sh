#!/bin/bash
# ncu captures of the traversal kernels on the DRAM-bound workload (C4, 16 Mi-state pool, 4 spp): launch list + full capture of trace / shadow.
TAG=${1:-r1}
WL=c4_instanced_10m_1080p_128spp
mkdir -p gpurun_out
export BTRC_ITER_LOG=1
PROF="python bench.py --workload $WL --steps 1 --warmup 3 --spp 4 --state-count 16777216 --no-cpu-baseline"
for K in trace_kernel shadow_kernel; do
  N=${K%%_kernel*}
  timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s 20 -c 3 -f -o gpurun_out/prof_c4_${N}_$TAG \
      $PROF > gpurun_out/ncu_c4_${N}_$TAG.log 2>&1
  grep "btrc iter" gpurun_out/ncu_c4_${N}_$TAG.log > gpurun_out/iters_c4_${N}_$TAG.log
  python scripts/ncu_summary.py traffic gpurun_out/prof_c4_${N}_$TAG.ncu-rep gpurun_out/iters_c4_${N}_$TAG.log 20 $N gpurun_out/traffic_$WL.json > /dev/null
done
cat gpurun_out/traffic_$WL.json | head -60
