"""Writes tests/golden/kat.json: known-answer vectors frozen from the oracle (oracle/liborc.so).

The reference holds no fixtures for RNG / hashing / alias tables / ray offsets (parity unpinned by the
reference), so these vectors pin the oracle against ITSELF over time (regressions) and give the CUDA kernels an
integer-exact target.  Run from the repo root:  python tests/golden/make_golden.py
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import front  # noqa: E402
from oracle.front import _ptr  # noqa: E402

f32 = np.float32
L = front.lib()
out = {"murmur": [], "pcg": []}
for px, py in [(0, 0), (1, 0), (0, 1), (511, 511), (1919, 1079), (3839, 2159), (123, 456)]:
    out["murmur"].append({"px": px, "py": py, "hash": str(L.orc_murmur_pixel(px, py))})
for px, py, sample in [(0, 0, 0), (0, 0, 1), (17, 3, 0), (511, 511, 15), (1919, 1079, 255), (3839, 2159, 1023)]:
    n = 16
    u = np.zeros(n, np.uint32)
    f = np.zeros(n, f32)
    si = np.zeros(2, np.uint64)
    L.orc_pcg_stream(px, py, sample, n, _ptr(u, ctypes.c_uint32), _ptr(f, ctypes.c_float), _ptr(si, ctypes.c_uint64))
    out["pcg"].append({"px": px, "py": py, "sample": sample, "state": str(int(si[0])), "inc": str(int(si[1])),
                       "u32": [int(x) for x in u], "f32_bits": [int(x) for x in f.view(np.uint32)]})
probs = (np.random.default_rng(3).random(37).astype(f32) ** 2).astype(f32)
probs[5] = 0
acc = np.zeros_like(probs)
ali = np.zeros(probs.size, np.uint32)
L.orc_alias_build(probs.size, _ptr(probs, ctypes.c_float), _ptr(acc, ctypes.c_float), _ptr(ali, ctypes.c_uint32))
out["alias"] = {"probs": [float(x) for x in probs], "accept_bits": [int(x) for x in acc.view(np.uint32)], "alias": [int(x) for x in ali]}
rng = np.random.default_rng(9)
pos = (rng.standard_normal((64, 3)) * np.array([1, 0.01, 100])).astype(f32)
pos[:4] = [[0, 0, 0], [0.03, -0.03, 0.031], [-5, 5, 1e-3], [1e4, -1e4, 3]]
nor = rng.standard_normal((64, 3)).astype(f32)
nor = (nor / np.linalg.norm(nor, axis=1, keepdims=True)).astype(f32)
d = rng.standard_normal((64, 3)).astype(f32)
res = np.zeros_like(pos)
L.orc_intersection_offset(64, _ptr(pos, ctypes.c_float), _ptr(nor, ctypes.c_float), _ptr(d, ctypes.c_float), _ptr(res, ctypes.c_float))
out["offset"] = {"pos": pos.tolist(), "nor": nor.tolist(), "dir": d.tolist(), "out_bits": [int(x) for x in res.view(np.uint32).reshape(-1)]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "kat.json"), "w"))
print("wrote kat.json")
