"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
f32 = np.float32

_libm = ctypes.CDLL("libm.so.6")
_libm.tanf.restype = ctypes.c_float
_libm.tanf.argtypes = [ctypes.c_float]


def bvh_benchmark_rays(width=1080, height=720, eye=(0, 0.9, 2.5), direction=(0, 0.001, -1), up=(0, 1, 0), fov=60.0):
    """The primary rays of lib/bvh's benchmark (lib/bvh/test/benchmark.cpp:120-148) in fp32, pixel (i, j) at row j.

    Returns o_med, d_t1 of shape (height*width, 4), index = width * j + i  (row 0 = bottom, as in the benchmark).
    """
    def dot(a, b):
        return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))

    def norm(v):
        inv = f32(1) / f32(np.sqrt(dot(v, v)))
        return np.array([v[0] * inv, v[1] * inv, v[2] * inv], f32)

    def cross(a, b):
        return np.array([f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]), f32(a[0] * b[1]) - f32(a[1] * b[0])], f32)

    eye = np.array(eye, f32)
    d = norm(np.array(direction, f32))
    image_u = norm(cross(d, np.array(up, f32)))
    image_v = norm(cross(image_u, d))
    image_w = f32(_libm.tanf(ctypes.c_float(float(f32(fov) * f32(3.14159265 * (1.0 / 180.0) * 0.5)))))
    ratio = f32(height) / f32(width)
    image_u = (image_u * image_w).astype(f32)
    image_v = ((image_v * image_w).astype(f32) * ratio).astype(f32)
    i = np.arange(width, dtype=f32)[None, :]
    j = np.arange(height, dtype=f32)[:, None]
    u = (f32(2) * (i + f32(0.5)) / f32(width) - f32(1)).astype(f32)
    v = (f32(2) * (j + f32(0.5)) / f32(height) - f32(1)).astype(f32)
    comp = []
    for k in range(3):
        comp.append(((image_u[k] * u).astype(f32) + (image_v[k] * v).astype(f32)).astype(f32) + d[k])
    x, y, z = (c.astype(f32) for c in comp)
    ln = np.sqrt(((x * x).astype(f32) + (y * y).astype(f32)).astype(f32) + (z * z).astype(f32)).astype(f32)
    inv = (f32(1) / ln).astype(f32)
    n = width * height
    d_t1 = np.zeros((n, 4), f32)
    d_t1[:, 0] = (x * inv).reshape(-1)
    d_t1[:, 1] = (y * inv).reshape(-1)
    d_t1[:, 2] = (z * inv).reshape(-1)
    d_t1[:, 3] = np.finfo(f32).max
    o_med = np.zeros((n, 4), f32)
    o_med[:, :3] = eye
    return o_med, d_t1


def normal_image_from_hits(tri_verts, prim, width, height):
    """Colour = |normalised triangle normal| of the hit primitive, quantised like benchmark.cpp:447-457 (PPM rows flipped)."""
    p0, p1, p2 = tri_verts[:, 0:3], tri_verts[:, 3:6], tri_verts[:, 6:9]
    e1 = (p0 - p1).astype(f32)
    e2 = (p2 - p0).astype(f32)
    n = np.stack([(e1[:, 1] * e2[:, 2]).astype(f32) - (e1[:, 2] * e2[:, 1]).astype(f32),
                  (e1[:, 2] * e2[:, 0]).astype(f32) - (e1[:, 0] * e2[:, 2]).astype(f32),
                  (e1[:, 0] * e2[:, 1]).astype(f32) - (e1[:, 1] * e2[:, 0]).astype(f32)], axis=1).astype(f32)
    ln = np.sqrt(((n[:, 0] * n[:, 0]).astype(f32) + (n[:, 1] * n[:, 1]).astype(f32)).astype(f32) + (n[:, 2] * n[:, 2]).astype(f32)).astype(f32)
    nn = np.abs(n * (f32(1) / ln)[:, None]).astype(f32)
    pix = np.zeros((height * width, 3), f32)
    hit = prim >= 0
    pix[hit] = nn[prim[hit]]
    img = np.clip((pix * f32(255)).astype(f32), 0, 255).astype(np.uint8).reshape(height, width, 3)
    return img[::-1]


def write_fan_cornell_scene(out_dir, width=64, height=64, spp=1):
    """The whole Cornell box as ONE mesh triangulated like lib/bvh's loader (fan), so primitive ids line up with lib/bvh."""
    import json

    from oracle import front

    os.makedirs(out_dir, exist_ok=True)
    tris = front.load_obj_fan(os.path.join(ROOT, "scenes", "assets", "cornell_box.obj"))
    with open(os.path.join(out_dir, "fan.obj"), "w") as f:
        for v in tris:
            f.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for t in range(tris.shape[0] // 3):
            f.write("f %d %d %d\n" % (3 * t + 1, 3 * t + 2, 3 * t + 3))
    root = {"width": width, "height": height,
            "camera": {"type": "pinhole", "eye": [0, 0.9, 2.5], "dst": [0, 0.9, 0], "up": [0, 1, 0], "fov_y": {"deg": 60}},
            "renderer": {"type": "wfpt", "spp": spp, "state_count": 100000},
            "scene": {"entities": [{"geometry": {"type": "triangle_mesh", "filename": "$env{scene_directory}/fan.obj"},
                                    "material": {"type": "diffuse", "color": [0.5, 0.5, 0.5]}}]}}
    path = os.path.join(out_dir, "scene.json")
    with open(path, "w") as f:
        json.dump(root, f)
    return path, tris


def rel_mse(a, b, eps=1e-2):
    """Mean per-pixel relative squared error (the usual relMSE): mean((a-b)^2 / (b^2 + eps))."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.mean((a - b) ** 2 / (b ** 2 + eps)))


def sort_by_key(arrs, key):
    order = np.lexsort(tuple(key[:, k] for k in range(key.shape[1] - 1, -1, -1))) if key.ndim == 2 else np.argsort(key, kind="stable")
    return {k: v[order] for k, v in arrs.items()}
