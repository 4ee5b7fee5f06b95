#!/bin/bash
# Round-2 final session: write-only HBM probe, then the whole GPU suite, smoke and the driver's two commands (scripts/gpu_r2k.sh).
TAG=${1:-r2final}
mkdir -p gpurun_out
python scripts/hbm_write_probe.py | tee gpurun_out/hbm_write_probe_$TAG.json
bash scripts/gpu_r2k.sh $TAG
