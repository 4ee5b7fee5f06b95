#!/bin/bash
# GPU parity tests only (+ optional memcheck of one small render).  Usage under gpurun: bash scripts/gpu_tests.sh <tag> [memcheck]
TAG=${1:-t}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 $PYTEST_ARGS 2>&1 | tail -60 | tee gpurun_out/pytest_gpu_$TAG.log
if [ "$2" = "memcheck" ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "
import tempfile
from btrc_b200 import api, scenes
d = tempfile.mkdtemp()
for p in (scenes.make_material_test(d + '/a', width=48, height=27, spp=2, rings=8, segments=12, tex_res=32, state_count=3000),
          scenes.make_medium_scene(d + '/b', width=32, height=32, spp=2, grid_res=12, state_count=3000)):
    s = api.Scene(p); s.render(); s.close()
print('memcheck render ok')
" 2>&1 | tail -25 | tee gpurun_out/memcheck_$TAG.log
fi
