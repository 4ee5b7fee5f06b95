#!/usr/bin/env python
"""Does material coherence inside a warp matter for the shade stage?  C2 as generated vs C2 with every sphere given the same
(first textured Disney) material: stage times of both (8 Mi pool, 32 spp)."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from btrc_b200 import api

tmp = tempfile.TemporaryDirectory()
base = bench.make_scene("c2_disney_1080p_256spp", tmp.name, spp=32)
root = json.load(open(base))
ents = root["scene"]["entities"]
disney = [e["material"] for e in ents if isinstance(e.get("material"), dict) and e["material"].get("type") == "disney"]
print("entities", len(ents), "disney materials", len(disney))
one = json.loads(json.dumps(root))
for e in one["scene"]["entities"]:
    if isinstance(e.get("material"), dict) and e["material"].get("type") in ("disney", "metal", "glass", "mirror") and "light" not in e:
        e["material"] = disney[1]
p1 = os.path.join(tmp.name, "one.json")
json.dump(one, open(p1, "w"))
for tag, path in (("as generated", base), ("one material", p1)):
    s = api.Scene(path, state_count=8388608)
    s.set_profiling(True)
    s.render()
    _, _, _, st = s.render()
    print("%-13s render %.1f ms  stages %s  n_closest %d n_shadow %d" % (tag, st["render_ms"], {k: round(v, 1) for k, v in st["stage_ms"].items()}, st["n_closest"], st["n_shadow"]), flush=True)
    s.close()
