#!/bin/bash
# A/B of environment knobs on the GPU box.  Usage: bash scripts/ab_env.sh <workload> <spp> <pool> "VAR=val VAR2=val" "..." ...
WL=$1; SPP=$2; POOL=$3; shift 3
mkdir -p gpurun_out
for E in "$@"; do
  echo "== $E"
  env $E python scripts/sweep_pool.py $WL $SPP $POOL
done 2>&1 | tee -a gpurun_out/ab_env.log
