#!/usr/bin/env python
"""One-off CPU check of the compressed 8-wide BVH at scale: the product's builder + traverse_bvh8 (host build of the same
sources, tests/host/host_traverse.cpp) against the oracle on a large instanced scene -- incoherent closest-hit and
occlusion rays, bit-exact.  The CPU suite runs the same check on small scenes (tests/test_traverse_host.py)."""
import ctypes
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from btrc_b200 import scenes  # noqa: E402
from oracle import front  # noqa: E402

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 50
L = ctypes.CDLL(os.path.join(ROOT, "tests", "host", "_libhosttrav_traverse_bvh2.so"))
L.ht8_build.restype = ctypes.c_void_p
L.ht8_build.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
L.ht8_info.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.ht8_closest.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
L.ht8_any.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
with tempfile.TemporaryDirectory() as d:
    path = scenes.make_instanced_scene(d, width=256, height=144, spp=1, n_instances=n_inst, n_lights=8, state_count=100000)
    t0 = time.time()
    orc = front.OracleScene(path)
    verts, ids = orc.triangles()
    verts = np.ascontiguousarray(verts, np.float32); ids = np.ascontiguousarray(ids, np.uint32)
    print("triangles", verts.shape[0], "oracle ready in %.1f s" % (time.time() - t0))
    t0 = time.time()
    h = L.ht8_build(verts.shape[0], verts.ctypes.data, ids.ctypes.data)
    print("wide8: %d nodes (%.1f MB against %.1f MB of binary nodes), %d levels, built in %.1f s" % (
        L.ht8_info(h, 0), L.ht8_info(h, 0) * 80 / 1e6, verts.shape[0] * 64 / 1e6 * 0.55, L.ht8_info(h, 1), time.time() - t0))
    rng = np.random.default_rng(3)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    n = 300000
    o = np.zeros((n, 4), np.float32); dd = np.zeros((n, 4), np.float32)
    o[:, :3] = rng.uniform(lo, hi, (n, 3))
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    dd[:, :3] = v
    dd[:, 3] = np.where(rng.random(n) < 0.5, np.finfo(np.float32).max, rng.uniform(0, np.linalg.norm(hi - lo), n)).astype(np.float32)
    flag, det = orc.closest_hit(o, dd)
    out_ids = np.zeros((n, 2), np.uint32); tuv = np.zeros((n, 3), np.float32); steps = np.zeros(2, np.uint64)
    L.ht8_closest(h, n, o.ctypes.data, dd.ctypes.data, out_ids.ctypes.data, tuv.ctypes.data, steps.ctypes.data)
    hit = ((flag >> 30) & 1).astype(bool)
    assert np.array_equal(hit, out_ids[:, 0] != 0xFFFFFFFF)
    assert np.array_equal(flag[hit] & 0x3FFFFFFF, out_ids[hit, 0]) and np.array_equal(det[hit, 1], out_ids[hit, 1])
    assert np.array_equal(det[hit][:, [0, 2, 3]], tuv[hit].view(np.uint32))
    occ = np.zeros(n, np.uint8)
    L.ht8_any(h, n, o.ctypes.data, dd.ctypes.data, occ.ctypes.data)
    assert np.array_equal(occ, orc.any_hit(o, dd))
    print("%d incoherent rays: %d hits, ids and t/u/v bit-identical to the oracle, occlusion identical; %.1f wide steps and %.1f triangle tests per ray"
          % (n, int(hit.sum()), steps[0] / n, steps[1] / n))
