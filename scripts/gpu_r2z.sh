#!/bin/bash
# Round-2 final profiler session (C2, final defaults): launch list + full captures of the four stage kernels, DRAM traffic per ray / state.

TAG=${1:-r2z}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
export BTRC_ITER_LOG=1
prof() {  # workload, short name, kernel regex (mangled base), stage key, skip, count, [extra env]
  WL=$1; N=$2; K=$3; STAGE=$4; SKIP=$5; CNT=$6; shift 6
  env "$@" timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/prof_${N}_$TAG \
      python bench.py --workload $WL --steps 1 --warmup 3 --spp ${SPP:-8} --no-cpu-baseline --no-workloads > gpurun_out/ncu_${N}_$TAG.log 2>&1
  grep "btrc iter" gpurun_out/ncu_${N}_$TAG.log > gpurun_out/iters_${N}_$TAG.log
  [ -n "$STAGE" ] && env "$@" python scripts/ncu_summary.py traffic gpurun_out/prof_${N}_$TAG.ncu-rep gpurun_out/iters_${N}_$TAG.log $SKIP $STAGE gpurun_out/traffic_$WL.json > /dev/null
  python scripts/ncu_summary.py full gpurun_out/prof_${N}_$TAG.ncu-rep gpurun_out/${TAG}_${N}_full.md > /dev/null 2>&1
  python scripts/ncu_hot_lines.py gpurun_out/prof_${N}_$TAG.ncu-rep 30 > gpurun_out/${TAG}_${N}_hot_lines.txt 2>/dev/null
  ls -la gpurun_out/prof_${N}_$TAG.ncu-rep 2>&1 | tail -1
  # the merged-back directory is capped at 64 MiB: only the two headline captures travel as .ncu-rep, the rest as summaries
  case $N in c2_trace|c2_shade) ;; *) rm -f gpurun_out/prof_${N}_$TAG.ncu-rep ;; esac
}
C2=c2_disney_1080p_256spp
cp profiles/traffic_$C2.json gpurun_out/traffic_$C2.json 2>/dev/null
SPP=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c2_$TAG.csv \
    python bench.py --workload $C2 --steps 1 --warmup 3 --spp 8 --no-cpu-baseline --no-workloads > gpurun_out/ncu_launch_c2_$TAG.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches_c2_$TAG.csv gpurun_out/${TAG}_c2_launches.md > /dev/null 2>&1
SPP=64 prof $C2 c2_trace trace_kernel trace 1 3
SPP=64 prof $C2 c2_shade shade_kernelILi2E shade 1 3
SPP=64 prof $C2 c2_shadow shadow_kernel shadow 1 3
SPP=64 prof $C2 c2_generate generate_kernel generate 0 2
rm -f gpurun_out/prof_c2_shade_$TAG.ncu-rep gpurun_out/launches_c2_$TAG.csv
ls -la gpurun_out | grep $TAG
