#!/usr/bin/env python
"""Freezes known-answer vectors of the oracle's volume-primitive medium (core/volume.cpp restatement) into
tests/golden/volumes_kat.json: ratio-tracking transmittance of fixed shadow rays through the two-box scene of
tests/test_oracle_volumes_cpu.py (streams keyed on the ray, so the values are deterministic)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from btrc_b200 import scenes  # noqa: E402
from oracle import front  # noqa: E402

f32 = np.float32
VOLS = [{"o": [-0.5, 0.5, -0.5], "x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1], "sigma_t": 2.0, "albedo": [0.8, 0.8, 0.8]},
        {"o": [0.0, 0.5, -0.5], "x": [1, 0.2, 0], "y": [0, 1, 0], "z": [0, 0.1, 1],
         "sigma_t": {"type": "array", "text": "$env{scene_directory}/density.txt"}, "albedo": [0.2, 0.9, 0.2]}]  # grid density: fractional ratios


def scene(d):
    path = scenes.make_volume_scene(d, width=16, height=16, spp=1, grid_res=8, state_count=1024)
    root = json.load(open(path))
    root["scene"]["volumes"] = VOLS
    json.dump(root, open(path, "w"))
    return path


def rays(o):
    rng = np.random.default_rng(99)
    n = 24
    a = rng.uniform(-0.9, 0.9, (n, 3)).astype(f32); a[:, 1] += f32(1.0)
    b = rng.uniform(-0.9, 0.9, (n, 3)).astype(f32); b[:, 1] += f32(1.0)
    o_med = np.zeros((n, 4), f32); d_t1 = np.zeros((n, 4), f32)
    o_med[:, :3] = a
    o_med[:, 3] = np.array([o.vol_prim_medium_id()], np.uint32).view(f32)[0]
    d_t1[:, :3] = b - a
    d_t1[:, 3] = 1.0
    return o_med, d_t1


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as d:
        o = front.OracleScene(scene(d))
        o_med, d_t1 = rays(o)
        tr = o.shadow_tr(o_med, d_t1)
    out = {"o_med_bits": o_med.view(np.uint32).tolist(), "d_t1_bits": d_t1.view(np.uint32).tolist(), "tr_bits": tr.view(np.uint32).tolist()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "volumes_kat.json"), "w"))
    print("wrote", len(out["tr_bits"]), "vectors; tr range", float(tr.min()), float(tr.max()))
