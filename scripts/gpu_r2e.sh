#!/bin/bash
# Round-2 multi-GPU session (N GPUs of one box): strong scaling of BASELINE config 5 with both partitions through the C ABI's
# own NCCL reduce, and the multi-rank CLI.
N=${1:-2}; TAG=${2:-r2e}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
nvidia-smi -L | head -8
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n${N}_tiles_$TAG.json 2> gpurun_out/bench_c5_n${N}_tiles_$TAG.err
run --steps 2 --warmup 3 --no-cpu-baseline --partition samples > gpurun_out/bench_c5_n${N}_samples_$TAG.json 2> gpurun_out/bench_c5_n${N}_samples_$TAG.err
python - <<PY
import json
for p in ("tiles", "samples"):
    try:
        d = json.load(open("gpurun_out/bench_c5_n${N}_%s_$TAG.json" % p))
        print(p, "value %.1f ms/step %.1f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["scaling"], d["multi_gpu"], d["run"]["partition"][:60])
    except Exception as e:
        print(p, "failed", e)
PY
tail -3 gpurun_out/bench_c5_n${N}_tiles_$TAG.err
# CLI: the same small scene on 1 and on N GPUs (sample split and tile split): images must agree
python - <<PY
import json, os
from btrc_b200 import scenes
for name, tile in (("u", False), ("t", True)):
    d = "/tmp/cli_" + name
    p = scenes.make_material_test(d, width=640, height=360, spp=32, rings=24, segments=32, tex_res=256, state_count=1048576, save="\$env{scene_directory}/out.hdr")
    r = json.load(open(p)); r["renderer"]["tile"] = tile; r["post_processors"] = [{"type": "save", "color": "\$env{scene_directory}/out.hdr"}]; json.dump(r, open(p, "w"))
PY
for name in u t; do
  ./btrc_b200/btrc /tmp/cli_$name/scene.json > gpurun_out/cli_${name}_n1_$TAG.log 2>&1; echo "cli $name n1 rc=$?"; mv /tmp/cli_$name/out.hdr /tmp/cli_$name/out_n1.hdr
  ./btrc_b200/btrc --ranks $N /tmp/cli_$name/scene.json > gpurun_out/cli_${name}_n${N}_$TAG.log 2>&1; echo "cli $name n$N rc=$?"; tail -3 gpurun_out/cli_${name}_n${N}_$TAG.log
done
python - <<PY
import numpy as np
from btrc_b200 import api
for name in ("u", "t"):
    a = api.load_image("/tmp/cli_%s/out_n1.hdr" % name).astype(np.float64); b = api.load_image("/tmp/cli_%s/out.hdr" % name).astype(np.float64)
    print("cli", name, "1 vs N GPUs: max abs diff %.3g of mean %.3g" % (np.abs(a - b).max(), a.mean()))
PY
