"""Array2D textures through the texture unit (`texture2d/array2d.cpp:18-30`, `utils/cuda/texture.cpp:63-105`,
`texture2d/description.cpp:28-80`): every address mode (clamp / wrap / mirror / border) x filter (linear / point), sRGB decode, one- and
four-channel 8-bit images, and the texture-graph operators (add, mul, transform), on the GPU against the oracle's texture-unit model.

One quad facing the camera with uv in [-1, 2]^2 carries the texture as its diffuse colour; the g-buffer albedo of a 1-spp render is
the texture value at every primary hit, so the comparison isolates texture evaluation.  Round 1 never exercised mirror / border on
the GPU; doing so (round 2, scripts/diag_textures.py) showed that the texture unit keeps the border colour IN THE TEXEL FORMAT
(quantised to 8 bits, ties down; one component only for one-channel textures) -- oracle/orc_texture.h now models that.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(d, color, name):
    from btrc_b200 import scenes

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
    p = os.path.join(d, "scene_%s.json" % name)
    json.dump(root, open(p, "w"))
    return p


def _variants(base, rough):
    tex = lambda name, **kw: dict({"type": "array", "filename": "$env{scene_directory}/" + name}, **kw)
    v = {}
    for mode in ("clamp", "wrap", "mirror", "border"):
        for filt in ("linear", "point"):
            kw = dict(address_mode_u=mode, address_mode_v=mode, filter=filt)
            if mode == "border":
                kw["border"] = [0.3]
            v["%s_%s" % (mode, filt)] = tex(base, **kw)
            v["%s_%s_1ch" % (mode, filt)] = tex(rough, **kw)
    v["border_rgb"] = tex(base, address_mode_u="border", address_mode_v="clamp", border=[0.1, 0.7, 0.9])
    v["border_u_wrap_v_point_1ch"] = tex(rough, address_mode_u="border", address_mode_v="wrap", filter="point", border=[0.5])
    v["srgb"] = tex(base, srgb_to_linear=True, address_mode_u="mirror", address_mode_v="wrap")
    v["transform_rotate"] = {"type": "transform", "transform": {"rotate": {"angle": {"deg": 30}}}, "tex": tex(base, address_mode_u="wrap", address_mode_v="wrap")}
    v["transform_swap_uv"] = {"type": "transform", "transform": {"swap_uv": True}, "tex": tex(base, address_mode_u="mirror", address_mode_v="border", border=[0.25])}
    v["add_const"] = {"type": "add", "lhs": tex(base), "rhs": [0.05, 0.1, 0.02]}
    v["sub_div"] = {"type": "div", "lhs": {"type": "sub", "lhs": tex(base, address_mode_u="wrap", address_mode_v="wrap"), "rhs": [0.02, 0.02, 0.02]}, "rhs": [1.5, 2.0, 2.5]}
    v["mul_two"] = {"type": "mul", "lhs": tex(base, address_mode_u="wrap", address_mode_v="wrap"), "rhs": tex(rough, address_mode_u="mirror", address_mode_v="mirror")}
    return v


def test_texture_unit_modes_and_graph_operators(tmp_path, built):
    from btrc_b200 import api, scenes
    from oracle import front

    d = str(tmp_path)
    base, rough, _ = scenes.write_test_textures(d, 32)
    report = []
    for name, color in _variants(base, rough).items():
        p = _scene(d, color, name)
        s, o = api.Scene(p), front.OracleScene(p)
        _, alb, _, _ = s.render()
        ref = o.render()["albedo"]
        s.close()
        a, b = alb[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
        hit = o.render(raw=True)["raw_weight"] > 0
        diff = np.abs(a - b).max(axis=-1)
        # 32 x 32 texels over ~50 pixels each: one 8-bit filter-weight step (1/256) of the largest texel contrast is ~4e-3;
        # a point-sampled pixel whose coordinate falls on a texel edge may take the neighbouring texel (counted, at most 3 per image)
        bad = diff > 1.5e-2
        report.append((name, float(np.median(diff)), float(np.percentile(diff, 99)), int(bad.sum())))
        assert np.median(diff) < 2.5e-3, (name, np.median(diff))
        assert bad.sum() <= 3, (name, int(bad.sum()), float(diff.max()))
    print("\nvariant                         median |d|   p99 |d|   pixels > 1.5e-2 (of 36864)")
    for r in report:
        print("%-30s  %.2e    %.2e   %d" % r)
