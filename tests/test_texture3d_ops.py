"""3-D texture operators "add" / "mul" (texture3d/binary.cpp): Texture3DBinaryOperatorCreator, its extrema by interval arithmetic
(the majorant of delta / ratio tracking) and its evaluation inside the heterogeneous medium and the volume primitives.

The anchor is an exact identity: with sigma_t = mul(grid, 2) the medium is the one whose grid file holds the doubled values -
scaling by a power of two commutes with every rounding of the trilinear filter, and max(grid) * 2 is the doubled grid's
maximum - so both scene files must give the same tracking decisions, bit for bit.  CPU leg: the oracle.  GPU leg: the CUDA
kernels (the instantiations that run 3-D texture programs against the plain ones), then the operator scene against the oracle.
"""
import json
import os

import numpy as np
import pytest

from tests import util
from tests.util import f32


def _write_grid(path, values):
    d, h, w = values.shape
    with open(path, "w") as f:
        f.write(f"float 1 {w} {h} {d}\n")
        f.write("\n".join("%.9g" % v for v in values.reshape(-1)))


def _scenes(tmp_path, res=16, width=48, height=48, spp=8):
    """Returns (operator scene, plain scene): the medium scene of BASELINE config 3 at test size, sigma_t = mul(add(grid, 0), 2)
    and albedo = mul([0.9 ...], 1) on one side, the doubled grid and the literal albedo on the other."""
    from btrc_b200 import scenes

    out = []
    rng = np.random.default_rng(21)
    grid = (rng.random((res, res, res)) ** 2 * 3.0).astype(f32)
    for name in ("ops", "plain"):
        d = str(tmp_path / name)
        path = scenes.make_medium_scene(d, width=width, height=height, spp=spp, grid_res=8, state_count=30000)
        _write_grid(os.path.join(d, "g1.txt"), grid)
        _write_grid(os.path.join(d, "g2.txt"), (grid * f32(2)).astype(f32))
        root = json.load(open(path))
        med = root["scene"]["entities"][-1]["inner_medium"]["medium"]
        if name == "ops":
            arr = {"type": "array", "text": "$env{scene_directory}/g1.txt"}
            med["sigma_t"] = {"type": "mul", "lhs": {"type": "add", "lhs": arr, "rhs": 0.0}, "rhs": {"type": "constant", "value": [2.0]}}
            med["albedo"] = {"type": "mul", "lhs": [0.9, 0.9, 0.9], "rhs": 1.0}
        else:
            med["sigma_t"] = {"type": "array", "text": "$env{scene_directory}/g2.txt"}
        json.dump(root, open(path, "w"))
        out.append(path)
    return out[0], out[1], grid


def test_oracle_texture3d_operators(tmp_path, built):
    from oracle import front

    ops, plain, grid = _scenes(tmp_path)
    a, b = front.OracleScene(ops), front.OracleScene(plain)
    ra, rb = a.render(), b.render()
    assert ra["stats"] == rb["stats"]
    assert np.array_equal(ra["color"], rb["color"])  # one path at a time, fixed splat order: bit-identical
    # extrema by interval arithmetic (texture3d/binary.cpp:95-137), through the oracle's C entry point
    L, h = a.L, a.h
    lo, hi = float(grid.min()), float(grid.max())
    g = L.orc_add_texture3d_array(h, 0)
    assert L.orc_texture3d_extremum(h, g, 0) == f32(lo) and L.orc_texture3d_extremum(h, g, 1) == f32(hi)
    neg = L.orc_add_texture3d_constant(h, -3.0, 0, 0)
    m = L.orc_add_texture3d_binary(h, 1, g, neg)        # grid * -3: the extrema swap
    assert L.orc_texture3d_extremum(h, m, 1) == f32(lo) * f32(-3) and L.orc_texture3d_extremum(h, m, 0) == f32(hi) * f32(-3)
    s = L.orc_add_texture3d_binary(h, 0, m, g)          # (grid * -3) + grid: interval bounds, not the tight ones
    assert L.orc_texture3d_extremum(h, s, 1) == f32(lo) * f32(-3) + f32(hi)
    assert L.orc_texture3d_extremum(h, s, 0) == f32(hi) * f32(-3) + f32(lo)
    with pytest.raises(ValueError, match="unknown texture3d type"):
        root = json.load(open(ops))
        root["scene"]["entities"][-1]["inner_medium"]["medium"]["sigma_t"]["type"] = "sub"   # only add and mul are registered (register.cpp:97-98)
        json.dump(root, open(ops, "w"))
        front.OracleScene(ops)


@pytest.mark.gpu
def test_gpu_texture3d_operators(tmp_path, built):
    from btrc_b200 import api
    from oracle import front

    ops, plain, _ = _scenes(tmp_path)
    a, b = api.Scene(ops), api.Scene(plain)
    assert a.info.has_medium == 1 and b.info.has_medium == 1
    ca, _, _, sa = a.render()
    cb, _, _, sb = b.render()
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert sa[k] == sb[k], (k, sa[k], sb[k])          # identical tracking decisions
    assert np.isclose(ca, cb, rtol=1e-4, atol=1e-6).mean() > 0.9999  # film atomics order is the only difference
    # the operator scene against the oracle: statistical match at equal spp and seed, as for the plain heterogeneous medium
    ref = front.OracleScene(ops).render()
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(sa[k] - ref["stats"][k]) <= 1e-2 * ref["stats"][k], (k, sa[k], ref["stats"][k])
    x, y = ca[..., :3], ref["color"][..., :3]
    assert abs(x.mean() - y.mean()) / y.mean() < 2e-2
    assert util.rel_mse(x, y) < 0.15
    a.close(); b.close()
    # volume primitives with operator densities: same identity
    from btrc_b200 import scenes
    paths = []
    for name in ("vops", "vplain"):
        d = str(tmp_path / name)
        path = scenes.make_volume_scene(d, width=48, height=36, spp=8, grid_res=8, state_count=20000)
        root = json.load(open(path))
        for vol in root["scene"]["volumes"]:
            if not isinstance(vol["sigma_t"], dict):
                base = float(vol["sigma_t"])
                vol["sigma_t"] = {"type": "mul", "lhs": base / 4, "rhs": 4.0} if name == "vops" else base
        json.dump(root, open(path, "w"))
        paths.append(path)
    a, b = api.Scene(paths[0]), api.Scene(paths[1])
    ca, _, _, sa = a.render()
    cb, _, _, sb = b.render()
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert sa[k] == sb[k], (k, sa[k], sb[k])
    assert np.isclose(ca, cb, rtol=1e-4, atol=1e-6).mean() > 0.9999
    a.close(); b.close()
    # unknown operator: the factory's message
    root = json.load(open(ops))
    root["scene"]["entities"][-1]["inner_medium"]["medium"]["sigma_t"]["type"] = "sub"
    json.dump(root, open(ops, "w"))
    with pytest.raises(api.BtrcError, match="unknown texture3d type: sub"):
        api.Scene(ops)
