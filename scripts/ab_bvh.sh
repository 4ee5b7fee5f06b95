#!/bin/bash
# A/B of the wide-BVH variants on the GPU box: collapse strategy (env BTRC_COLLAPSE) x warp-vote traversal (two builds of the library).
mkdir -p gpurun_out
for LIB in libbtrc_b200.so libbtrc_b200_nc.so; do
  for M in dp count area; do
    echo "== $LIB collapse=$M"
    BTRC_B200_LIB=$PWD/btrc_b200/$LIB BTRC_COLLAPSE=$M python scripts/sweep_pool.py c2_disney_1080p_256spp 32 8388608
  done
done 2>&1 | tee gpurun_out/ab_bvh_c2.log
for M in dp count; do
  echo "== libbtrc_b200.so collapse=$M"
  BTRC_COLLAPSE=$M python scripts/sweep_pool.py c4_instanced_10m_1080p_128spp 8 16777216
done 2>&1 | tee gpurun_out/ab_bvh_c4.log
