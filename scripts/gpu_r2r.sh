#!/bin/bash
# Round-2 A/B: order inside generation blocks / partition tiles (pixel-major vs 32-sample chunks), strips vs 2-D tiles at N GPUs.
N=${1:-2}; TAG=${2:-r2r}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_config_size.py -m gpu -q -x -k "ragged or c5" 2>&1 | tail -5
for blk in 64x32 64x32c 32x32c 64x16c; do
  BTRC_GEN_BLOCK=$blk python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_c2_${blk}_$TAG.json 2> gpurun_out/bench_c2_${blk}_$TAG.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_c2_${blk}_$TAG.json"))
print("C2 block $blk value %.1f" % d["value"], {k: round(v["ms_total"], 1) for k, v in d["roofline_stages"].items()})
PY
done
for t2d in 1 0; do
  BTRC_TILES_2D=$t2d python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n${N}_t2d${t2d}_$TAG.json 2> gpurun_out/bench_c5_n${N}_t2d${t2d}_$TAG.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_c5_n${N}_t2d${t2d}_$TAG.json"))
print("C5 N=$N tiles2d=$t2d value %.1f ms/step %.1f" % (d["value"], d["ms_per_step"]), d["multi_gpu"])
PY
done
