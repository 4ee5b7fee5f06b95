"""GPU diagnostic: which texture feature disagrees with the oracle?  One quad facing the camera with uv in [-1, 2]^2, the texture
as the diffuse colour, 1 spp; the g-buffer albedo is the texture value at every primary hit.  Prints, per variant, the fraction of
pixels whose albedo differs by more than 2e-2 and where they lie in uv."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from btrc_b200 import api, scenes  # noqa: E402
from oracle import front  # noqa: E402


def scene(d, color):
    gv, gf, guv = scenes.quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0))
    scenes.write_obj(os.path.join(d, "q.obj"), gv, gf, [(u * 3 - 1, w * 3 - 1) for u, w in guv])
    lv, lf, luv = scenes.quad((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1))
    scenes.write_obj(os.path.join(d, "l.obj"), lv, lf, luv)
    geo = lambda n: {"type": "triangle_mesh", "filename": "$env{scene_directory}/%s.obj" % n}
    root = {"width": 192, "height": 192,
            "camera": {"type": "pinhole", "eye": [0, 0, 3], "dst": [0, 0, 0], "up": [0, 1, 0], "fov_y": {"deg": 40}},
            "renderer": {"type": "wfpt", "spp": 1, "state_count": 100000, "min_depth": 1, "max_depth": 1, "albedo": True, "filter": {"type": "box"}},
            "scene": {"entities": [{"geometry": geo("q"), "material": {"type": "diffuse", "color": color}},
                                   {"geometry": geo("l"), "material": {"type": "diffuse", "color": [0, 0, 0]}, "local_to_world": [{"translate": [0, 5, 0]}],
                                    "light": {"type": "mesh_light", "intensity": [1, 1, 1]}}]}}
    p = os.path.join(d, "scene.json")
    json.dump(root, open(p, "w"))
    return p


def main():
    d = tempfile.mkdtemp()
    base, rough, nmap = scenes.write_test_textures(d, 32)
    tex = lambda name, **kw: dict({"type": "array", "filename": "$env{scene_directory}/" + name}, **kw)
    variants = {
        "clamp_linear": tex(base),
        "wrap_linear": tex(base, address_mode_u="wrap", address_mode_v="wrap"),
        "mirror_linear": tex(base, address_mode_u="mirror", address_mode_v="mirror"),
        "border_linear": tex(base, address_mode_u="border", address_mode_v="border", border=[0.5]),
        "clamp_point": tex(base, filter="point"),
        "wrap_point": tex(base, filter="point", address_mode_u="wrap", address_mode_v="wrap"),
        "mirror_point": tex(base, filter="point", address_mode_u="mirror", address_mode_v="mirror"),
        "border_point": tex(base, filter="point", address_mode_u="border", address_mode_v="border", border=[0.5]),
        "border_u_wrap_v_point": tex(rough, address_mode_u="border", address_mode_v="wrap", filter="point", border=[0.5]),
        "transform_rotate": {"type": "transform", "transform": {"rotate": {"angle": {"deg": 30}}}, "tex": tex(base, address_mode_u="wrap", address_mode_v="wrap")},
        "add_const": {"type": "add", "lhs": tex(base), "rhs": [0.05, 0.1, 0.02]},
        "mul_two": {"type": "mul", "lhs": tex(base, address_mode_u="wrap", address_mode_v="wrap"), "rhs": tex(rough, address_mode_u="wrap", address_mode_v="wrap")},
        "srgb": tex(base, srgb_to_linear=True),
    }
    for name, color in variants.items():
        p = scene(d, color)
        s, o = api.Scene(p), front.OracleScene(p)
        _, alb, _, _ = s.render()
        ref = o.render()["albedo"]
        a, b = alb[..., :3], ref[..., :3]
        diff = np.abs(a - b).max(axis=-1)
        hit = b.max(axis=-1) > 0
        bad = (diff > 2e-2) & hit
        ys, xs = np.nonzero(bad)
        where = "" if not bad.any() else " bad pixels x in [%d, %d], y in [%d, %d]" % (xs.min(), xs.max(), ys.min(), ys.max())
        print("%-24s hit %.2f  median |d| %.2e  p99 %.2e  frac > 2e-2: %.4f%s" % (name, hit.mean(), np.median(diff[hit]), np.percentile(diff[hit], 99), bad.sum() / max(hit.sum(), 1), where))
        if bad.any():
            k = np.argmax(diff * hit)
            y, x = divmod(k, a.shape[1])
            print("      worst at (%d, %d): gpu %s oracle %s" % (x, y, a[y, x], b[y, x]))
        s.close()


if __name__ == "__main__":
    main()
