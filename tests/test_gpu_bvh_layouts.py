"""GPU parity of the alternative BVH layouts the renderer node can ask for ("bvh_layout": "wide8" | "instanced"; default "flat").

The reference hands its triangles to OptiX (utils/optix/context.cpp:158-294) and has no layout of its own to match: what must hold is
that every layout returns the SAME hit -- bit-identical (instance, primitive, t, u, v) and occlusion against the oracle on the same ray
batches (primary, bounces down to depth 16, shadow rays), the same render statistics, and the same film as the default layout up to the
order of the float additions inside a pixel.  "wide8" is the compressed 8-wide tree of BASELINE's north_star (80-byte nodes, 8-bit
child boxes); "instanced" the two-level tree.  Both are measured slower than the flat binary tree on B200 (profiles/r2v_ab_wide_bvh8.md,
r2c_ab_bvh_layout.md), hence selectable, not default.
"""
import json

import numpy as np
import pytest

from tests.test_gpu_config_size import _raw_film
from tests.test_gpu_instanced import _bounce_states

pytestmark = pytest.mark.gpu

W, H = 128, 72


def _with_layout(path, layout):
    root = json.load(open(path))
    root["renderer"]["bvh_layout"] = layout
    out = path.replace(".json", "_%s.json" % layout)
    json.dump(root, open(out, "w"))
    return out


def _scene(kind, d):
    from btrc_b200 import scenes

    if kind == "instanced":
        return scenes.make_instanced_scene(d, width=W, height=H, spp=4, rings=21, segments=40, n_instances=10, n_lights=4, state_count=40000)
    if kind == "materials":
        return scenes.make_material_test(d, width=W, height=H, spp=4, rings=24, segments=32, state_count=40000)
    return scenes.make_cornell(d, width=96, height=96, spp=4, state_count=20000)


@pytest.mark.parametrize("layout", ["wide8", "instanced"])
@pytest.mark.parametrize("kind", ["instanced", "materials", "cornell"])
def test_layout_returns_the_same_hits_and_film(tmp_path, built, kind, layout):
    from btrc_b200 import api
    from oracle import front

    path = _scene(kind, str(tmp_path))
    o = front.OracleScene(path)
    flat = api.Scene(path)
    s = api.Scene(_with_layout(path, layout))
    assert bool(s.info.bvh_wide) == (layout == "wide8") and bool(s.info.bvh_instanced) == (layout == "instanced")
    assert not flat.info.bvh_wide and not flat.info.bvh_instanced
    if layout == "wide8" and kind != "cornell":
        assert s.info.bvh_bytes < flat.info.bvh_bytes   # the compact layout (36 triangles: one node either way)
    total = 0
    for st, flag, detail, sh in _bounce_states(o, 0, o.width * o.height * 2, 16):
        f1, d1 = s.closest_hit(st["o_med"], st["d_t1"])
        assert np.array_equal(f1, flag)
        hit = ((flag >> 30) & 1).astype(bool)
        assert np.array_equal(d1[hit], detail[hit])
        has_sh = (sh["status"] & 2) != 0
        so, sd = np.ascontiguousarray(sh["sh_o_med"][has_sh]), np.ascontiguousarray(sh["sh_d_t1"][has_sh])
        assert np.array_equal(s.any_hit(so, sd), o.any_hit(so, sd))
        total += st["o_med"].shape[0] + so.shape[0]
    assert total > 30000

    ra, wa, sa = _raw_film(flat)
    rb, wb, sb = _raw_film(s)
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"):
        assert sa[k] == sb[k], (k, sa[k], sb[k])      # the same hits => the same paths, event for event
    assert np.array_equal(wa, wb)
    assert np.abs(ra - rb).max() <= 1e-5 * max(1.0, float(ra.max()))
    s.close(); flat.close()
