"""GPU diagnostic: how does the texture unit quantise the border colour of a UNorm8 texture?"""
import json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from btrc_b200 import api, scenes
from scripts.diag_textures import scene

d = tempfile.mkdtemp()
base, rough, nmap = scenes.write_test_textures(d, 32)
for b in (0.1, 0.2, 0.25, 0.3, 0.5, 0.7, 0.9, 0.998, 1.0, 1.5, -0.2):
    for name in (base, rough):
        p = scene(d, {"type": "array", "filename": "$env{scene_directory}/" + name, "address_mode_u": "border", "address_mode_v": "border", "filter": "point", "border": [b]})
        s = api.Scene(p)
        _, alb, _, _ = s.render()
        v = alb[30, 30, :3]   # a pixel in the border region (uv < 0)
        print("border %.3f %-12s -> %s   x*255 = %.4f" % (b, name, v, v[0] * 255))
        s.close()
