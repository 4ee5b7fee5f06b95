#!/bin/bash
# Round-2 A/B: hot counters of the wave control block on separate 128-byte lines (BTRC_CTL_PAD), static + dynamic batch hand-out
# (BTRC_STATIC_SHARE) -- full variant libraries under btrc_b200/variants/ (scripts/build_variant.py), selected with BTRC_B200_LIB.
TAG=${1:-r2ab}; shift
VARIANTS=${@:-default ctlpad static80 static80pad default}
mkdir -p gpurun_out
for V in $VARIANTS; do
  L=btrc_b200/libbtrc_b200.so; [ $V != default ] && L=btrc_b200/variants/libbtrc_$V.so
  [ -f $L ] || { echo "$V: no library"; continue; }
  for WL in c2_disney_1080p_256spp:64:5 c4_instanced_10m_1080p_128spp:16:2 c3_medium_1080p_128spp:64:3; do
    IFS=: read W SPP STEPS <<< "$WL"; S=${W%%_*}
    BTRC_B200_LIB=$PWD/$L python bench.py --workload $W --spp $SPP --steps $STEPS --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_${S}_${V}_$TAG.json 2> gpurun_out/bench_${S}_${V}_$TAG.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${S}_${V}_$TAG.json"))
    print("$S $V", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
except Exception as e:
    print("$S $V failed", e, open("gpurun_out/bench_${S}_${V}_$TAG.err").read()[-300:])
PY
  done
done
