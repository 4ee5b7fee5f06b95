#!/bin/bash
# ncu --set full capture of selected kernels on the default workload at 8 spp.  Usage: bash scripts/gpu_ncu.sh <tag> <kernel regex> [...]
TAG=$1; shift
mkdir -p gpurun_out
PROF="python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline"
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-2} -f -o gpurun_out/prof_${K}_$TAG $PROF > gpurun_out/ncu_${K}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${K}_$TAG.log
done
