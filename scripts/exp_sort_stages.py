#!/usr/bin/env python
"""Stage times of one workload with and without the binned ray sort (BTRC_SORT_RAYS), one process per setting."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "c2_disney_1080p_256spp"
spp = sys.argv[2] if len(sys.argv) > 2 else "32"
pool = sys.argv[3] if len(sys.argv) > 3 else "8388608"
code = r'''
import sys, tempfile
sys.path.insert(0, %r)
import bench
from btrc_b200 import api
tmp = tempfile.TemporaryDirectory()
path = bench.make_scene(%r, tmp.name, spp=%s)
s = api.Scene(path, state_count=%s)
s.set_profiling(True)
s.render()
_, _, _, st = s.render()
print("render %%.1f ms  stages %%s" %% (st["render_ms"], {k: round(v, 1) for k, v in st["stage_ms"].items()}))
''' % (ROOT, wl, spp, pool)
for v in ("0", "1", "2"):
    env = dict(os.environ, BTRC_SORT_RAYS=v)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("BTRC_SORT_RAYS=%s" % v, out.stdout.strip(), out.stderr.strip()[-300:])
