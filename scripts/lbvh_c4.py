#!/usr/bin/env python
"""Device LBVH vs host SAH on C4 (10 M triangles): build time and render time (16 Mi pool, 8 spp)."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from btrc_b200 import api

tmp = tempfile.TemporaryDirectory()
base = bench.make_scene("c4_instanced_10m_1080p_128spp", tmp.name, spp=8)
root = json.load(open(base))
root["renderer"]["bvh_builder"] = "lbvh"
lpath = os.path.join(tmp.name, "scene_lbvh.json")
json.dump(root, open(lpath, "w"))
for tag, path in (("sah", base), ("lbvh", lpath)):
    t0 = time.time()
    s = api.Scene(path, state_count=16777216)
    t1 = time.time()
    s.set_profiling(False)
    color, _, _, st = s.render()
    t2 = time.time()
    ms = []
    for k in range(3):
        color, _, _, st = s.render()
        ms.append(st["render_ms"])
    print("%-5s load+commit %.1f s  first render (incl. upload%s) %.1f s  bvh_build %.3f s  render %.1f ms  n_closest %d n_hit %d" %
          (tag, t1 - t0, " + device build" if tag == "lbvh" else "", t2 - t1, s.info.bvh_build_seconds, min(ms), st["n_closest"], st["n_hit"]), flush=True)
    s.close()
