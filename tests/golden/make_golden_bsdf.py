"""Writes tests/golden/bsdf_kat.json: known-answer vectors of every material's sample / eval / pdf frozen from the oracle.

Regression pins only (the reference ships no fixtures for its BSDFs: SURVEY 8c).  Run from the repo root:
    python tests/golden/make_golden_bsdf.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import front  # noqa: E402
from tests.test_oracle_materials_cpu import MATERIALS, material_scene  # noqa: E402

f32 = np.float32
rng = np.random.default_rng(2024)
n = 12
wo = rng.standard_normal((n, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True); wo[:8, 2] = np.abs(wo[:8, 2])
wi = rng.standard_normal((n, 3)); wi /= np.linalg.norm(wi, axis=1, keepdims=True)
sam = rng.random((n, 3))
wo, wi, sam = wo.astype(f32), wi.astype(f32), sam.astype(f32)
interp_z, uv = [0.08, -0.05, 1.0], [0.3, 0.6]
o = front.OracleScene(material_scene(tempfile.mkdtemp()))
out = {"wo": wo.tolist(), "wi": wi.tolist(), "sam": sam.tolist(), "interp_z": interp_z, "uv": uv, "materials": {}}
for mid, name in enumerate(MATERIALS):
    r = o.bsdf_probe(mid, interp_z, uv, wo, wi, sam)
    out["materials"][name] = {"sample": r["sample"].tolist(), "eval": r["eval"].tolist(), "pdf": r["pdf"].tolist()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bsdf_kat.json"), "w"))
print("wrote bsdf_kat.json")
