#!/bin/bash
# Round-2 last capture: the closest-hit kernel on C2 after the wave-control counters moved to separate lines (compare r2z_c2_trace_hot_lines.txt).
TAG=r2last
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:trace_kernel -s 1 -c 1 -f -o gpurun_out/prof_c2_trace_$TAG \
    python bench.py --workload c2_disney_1080p_256spp --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-workloads > gpurun_out/ncu_c2_trace_$TAG.log 2>&1
python scripts/ncu_hot_lines.py gpurun_out/prof_c2_trace_$TAG.ncu-rep 30 > gpurun_out/${TAG}_c2_trace_hot_lines.txt 2>/dev/null
python scripts/ncu_summary.py full gpurun_out/prof_c2_trace_$TAG.ncu-rep gpurun_out/${TAG}_c2_trace_full.md > /dev/null 2>&1
rm -f gpurun_out/prof_c2_trace_$TAG.ncu-rep
head -12 gpurun_out/${TAG}_c2_trace_hot_lines.txt | cut -c1-150
