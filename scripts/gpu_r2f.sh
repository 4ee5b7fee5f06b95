#!/bin/bash
# Round-2 8-GPU check: BASELINE config 5 (4K, 1024 spp, strong scaling) exactly as the driver launches it.
N=${1:-8}; TAG=${2:-r2f}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n${N}_$TAG.json 2> gpurun_out/bench_c5_n${N}_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_c5_n${N}_$TAG.json"))
print("value %.1f ms/step %.1f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["scaling"], d["multi_gpu"], d["clocks"])
PY
tail -3 gpurun_out/bench_c5_n${N}_$TAG.err
