#!/bin/bash
# A/B of library builds on the GPU box.  Usage: bash scripts/ab_libs.sh <workload> <spp> <pool> lib1.so lib2.so ...
WL=$1; SPP=$2; POOL=$3; shift 3
mkdir -p gpurun_out
for LIB in "$@"; do
  echo "== $LIB"
  BTRC_B200_LIB=$PWD/btrc_b200/$LIB python scripts/sweep_pool.py $WL $SPP $POOL
done 2>&1 | tee -a gpurun_out/ab_libs.log
