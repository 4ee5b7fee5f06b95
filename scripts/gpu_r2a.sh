#!/bin/bash
# Round-2 first GPU session: counters for the kernels round 1 left without evidence (C3 medium/shade/shadow/trace, generate),
# and driver-style bench lines for C1 / C3 / C4 with the clocks record.
TAG=${1:-r2a}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
export BTRC_ITER_LOG=1
PROF="python bench.py --workload c3_medium_1080p_128spp --steps 1 --warmup 3 --spp 8 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c3_$TAG.csv $PROF > gpurun_out/ncu_launch_c3_$TAG.log 2>&1
for K in medium_kernel generate_kernel shade_kernel shadow_kernel trace_kernel; do
  N=${K%%_kernel*}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 3 -f -o gpurun_out/prof_c3_${N}_$TAG $PROF > gpurun_out/ncu_c3_${N}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_c3_${N}_$TAG.log
done
bash scripts/gpu_workloads.sh $TAG
ls -la gpurun_out | tail -30
