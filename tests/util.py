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


def shadow_ray_owners(shadow_pixel, shadow_o, st):
    """Shadow rays come back compacted in any order.  Each belongs to exactly one input state: the one of its pixel whose ray
    (o + s * d) passes closest to the shadow origin (a surface NEE ray starts at the offset hit point, a medium NEE ray at
    the scatter position on the ray).  Returns the owning state index per shadow ray."""
    from collections import defaultdict

    bypix = defaultdict(list)
    for i, (px, py) in enumerate(st["pixel"]):
        bypix[(int(px), int(py))].append(i)
    o = st["o_med"][:, :3].astype(np.float64)
    d = st["d_t1"][:, :3].astype(np.float64)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    owner = np.full(shadow_pixel.shape[0], -1, np.int64)
    for j, (px, py) in enumerate(shadow_pixel):
        c = np.array(bypix[(int(px), int(py))], np.int64)
        w = shadow_o[j, :3].astype(np.float64)[None, :] - o[c]
        dist = np.linalg.norm(w - (w * d[c]).sum(axis=1, keepdims=True) * d[c], axis=1)
        owner[j] = c[int(np.argmin(dist))]
    return owner


def compare_shadow_rays(shd, st, has_sh, ref_o, ref_d, ref_beta, rtol, atol_dir=5e-5, max_flip=0.0, max_outlier=0.0, negligible=1e-4):
    """Compares compacted shadow rays with the oracle's per-state shadow outputs (ref_* indexed by state, has_sh = mask).

    A ray may exist on one side only when the BSDF value towards the light sample is zero on the other side; for a sample in
    the surface's own plane (a path sitting on the emitter) that is the sign of a rounding residue and the ray carries
    |cos| ~ 1e-7.  Such rays must be negligible (< `negligible` x the largest beta_li); `max_flip` bounds one-sided rays that
    do carry energy (decision flips of textured / glossy materials), `max_outlier` the fraction outside the tolerances."""
    n = st["pixel"].shape[0]
    owner = shadow_ray_owners(shd["pixel"], shd["o_med"], st)
    cnt = np.bincount(owner, minlength=n)
    assert (cnt <= 1).all(), "two shadow rays matched one state"
    scale = max(float(np.abs(ref_beta[has_sh, :3]).max()), 1e-6) if has_sh.any() else 1.0
    got = cnt > 0
    ray_of = np.full(n, -1, np.int64)
    ray_of[owner] = np.arange(owner.shape[0])
    only_gpu = np.where(got & ~has_sh)[0]
    only_ref = np.where(~got & has_sh)[0]
    energetic = int((np.abs(shd["beta_li"][ray_of[only_gpu], :3]).max(axis=1) > negligible * scale).sum()) if only_gpu.size else 0
    energetic += int((np.abs(ref_beta[only_ref, :3]).max(axis=1) > negligible * scale).sum()) if only_ref.size else 0
    assert only_gpu.size + only_ref.size <= max(0.01, max_flip) * max(int(has_sh.sum()), 100), (only_gpu.size, only_ref.size)
    assert energetic <= max_flip * max(int(has_sh.sum()), 100), ("one-sided shadow rays carry energy", energetic)
    both = np.where(got & has_sh)[0]
    if both.size == 0:
        return
    j = ray_of[both]

    def frac(a, b, r, t):
        return float(np.isclose(a, b, rtol=r, atol=t).all(axis=-1).mean())

    assert np.array_equal(shd["pixel"][j], st["pixel"][both])
    assert (shd["o_med"][j, 3].view(np.uint32) == ref_o[both, 3].view(np.uint32)).mean() >= 1 - max_outlier  # medium id
    assert frac(shd["o_med"][j, :3], ref_o[both, :3], 0, atol_dir) >= 1 - max_outlier
    assert frac(shd["d_t1"][j], ref_d[both], 0, atol_dir) >= 1 - max_outlier
    assert frac(shd["beta_li"][j, :3], ref_beta[both, :3], rtol, 1e-6 * scale) >= 1 - max_outlier
