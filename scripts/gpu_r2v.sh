#!/bin/bash
# Round-2 session v: the compressed 8-wide BVH layout as a product option -- GPU parity, A/B against the flat binary tree on C2 and C4,
# ncu launch list + full capture of its trace kernel.
TAG=${1:-r2v}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_bvh_layouts.py -m gpu -q -x 2>&1 | tail -5
for L in flat wide8; do
  BTRC_BVH_LAYOUT=$L python bench.py --workload c2_disney_1080p_256spp --spp 64 --steps 5 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_c2_${L}_$TAG.json 2> gpurun_out/bench_c2_${L}_$TAG.err
  BTRC_BVH_LAYOUT=$L python bench.py --workload c4_instanced_10m_1080p_128spp --spp 16 --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/bench_c4_${L}_$TAG.json 2> gpurun_out/bench_c4_${L}_$TAG.err
  python - <<PY
import json
for wl in ("c2", "c4"):
    try:
        d = json.load(open("gpurun_out/bench_%s_${L}_$TAG.json" % wl))
        print(wl, "$L", "value %.1f ms/step %.2f" % (d["value"], d["ms_per_step"]), d["run"]["bvh_layout"], {k: round(v["ms_per_launch"], 3) for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(wl, "$L failed", e, open("gpurun_out/bench_%s_${L}_$TAG.err" % wl).read()[-400:])
PY
done
for WL in c2_disney_1080p_256spp c4_instanced_10m_1080p_128spp; do
  S=${WL%%_*}
  BTRC_BVH_LAYOUT=wide8 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:trace_kernel -s 1 -c 3 -f -o gpurun_out/prof_${S}_trace_wide8_$TAG \
    python bench.py --workload $WL --spp 8 --steps 1 --warmup 3 --no-cpu-baseline --no-workloads > gpurun_out/ncu_${S}_wide8_$TAG.log 2>&1
  python scripts/ncu_summary.py full gpurun_out/prof_${S}_trace_wide8_$TAG.ncu-rep gpurun_out/${TAG}_${S}_trace_wide8_full.md > /dev/null 2>&1
  rm -f gpurun_out/prof_${S}_trace_wide8_$TAG.ncu-rep
  ls -la gpurun_out/${TAG}_${S}_trace_wide8_full.md
done
tail -3 gpurun_out/ncu_c4_wide8_$TAG.log
