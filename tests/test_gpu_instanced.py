"""GPU parity for the instanced-mesh scene (BASELINE config 4: 250 x 40,000 triangles, 64 mesh lights, max depth 16).

Reduced size (10 instances of an 1,600-triangle mesh, 4 lights): stages and the whole render against the oracle.
Full size (10,000,140 triangles): closest hit / occlusion on incoherent ray batches bit-exact against the oracle, and the
size-independent properties of a low-spp render (weights == spp, linearity over sample ranges).
"""
import numpy as np
import pytest

from tests import util
from tests.util import f32

pytestmark = pytest.mark.gpu

W4, H4 = 128, 72


def _bounce_states(o, first, n, bounces):
    st = o.generate(first, n)
    out = []
    for _ in range(bounces):
        flag, detail = o.closest_hit(st["o_med"], st["d_t1"])
        sh = o.shade(st, flag, detail)
        out.append((st, flag, detail, sh))
        keep = (sh["status"] & 1) != 0
        st = {k: np.ascontiguousarray(v[keep]) for k, v in sh["next"].items()}
        if st["o_med"].shape[0] == 0:
            break
    return out


@pytest.fixture(scope="module")
def pair_small(tmp_path_factory, built):
    from btrc_b200 import api, scenes
    from oracle import front

    d = tmp_path_factory.mktemp("c4_small")
    path = scenes.make_instanced_scene(str(d), width=W4, height=H4, spp=8, rings=21, segments=40, n_instances=10, n_lights=4, state_count=40000)
    s = api.Scene(path)
    o = front.OracleScene(path)
    yield s, o
    s.close()


def test_instanced_tables(pair_small):
    s, o = pair_small
    i = s.info
    assert i.n_instances == 6 + 10 + 4 and i.n_lights == 4
    assert i.n_triangles == o.num_triangles() == 12 + 10 * 1600 + 4 * 2
    assert i.max_depth == 16


def test_instanced_closest_and_any_hit(pair_small):
    """Shared geometry under 10 different rigid transforms + scales: identical ray batches -> bit-identical hits, down to depth 16."""
    s, o = pair_small
    total = 0
    for st, flag, detail, sh in _bounce_states(o, 0, W4 * H4 * 2, 16):
        f1, d1 = s.closest_hit(st["o_med"], st["d_t1"])
        assert np.array_equal(f1, flag)
        hit = ((flag >> 30) & 1).astype(bool)
        assert np.array_equal(d1[hit], detail[hit])
        has_sh = (sh["status"] & 2) != 0
        so, sd = np.ascontiguousarray(sh["sh_o_med"][has_sh]), np.ascontiguousarray(sh["sh_d_t1"][has_sh])
        assert np.array_equal(s.any_hit(so, sd), o.any_hit(so, sd))
        total += st["o_med"].shape[0] + so.shape[0]
    assert total > 100000


def test_instanced_render_matches_oracle(pair_small):
    s, o = pair_small
    color, _, _, stats = s.render()
    ref = o.render()
    rs = ref["stats"]
    assert stats["n_gen"] == rs["n_gen"] == W4 * H4 * 8
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(stats[k] - rs[k]) <= max(4, 2e-3 * rs[k]), (k, stats[k], rs[k])
    a, b = color[..., :3], ref["color"][..., :3]
    assert np.isfinite(a).all()
    # diffuse + metal (fp32 fast-math shading): same seeds and spp, paths differ only where a discrete decision flips
    assert util.rel_mse(a, b) < 1e-2
    assert abs(a.mean() - b.mean()) / b.mean() < 5e-3


def test_instanced_full_size_10m(tmp_path_factory, built):
    """BASELINE config 4 at full size: 10,000,140 world-space triangles in one BVH."""
    from btrc_b200 import api, scenes
    from oracle import front

    d = tmp_path_factory.mktemp("c4_full")
    W, H, SPP = 480, 270, 2
    path = scenes.make_instanced_scene(str(d), width=W, height=H, spp=SPP, state_count=300000)
    s = api.Scene(path)
    o = front.OracleScene(path)
    assert s.info.n_triangles == o.num_triangles() == 10000140 and s.info.n_lights == 64
    total = 0
    for st, flag, detail, sh in _bounce_states(o, 0, 60000, 6):
        f1, d1 = s.closest_hit(st["o_med"], st["d_t1"])
        assert np.array_equal(f1, flag)
        hit = ((flag >> 30) & 1).astype(bool)
        assert np.array_equal(d1[hit], detail[hit])
        has_sh = (sh["status"] & 2) != 0
        so, sd = np.ascontiguousarray(sh["sh_o_med"][has_sh]), np.ascontiguousarray(sh["sh_d_t1"][has_sh])
        assert np.array_equal(s.any_hit(so, sd), o.any_hit(so, sd))
        total += st["o_med"].shape[0] + so.shape[0]
    assert total > 300000
    # size-independent properties of the render
    npix = W * H
    color, _, _, stats = s.render()
    assert stats["n_gen"] == npix * SPP and stats["n_closest"] == stats["n_gen"] + stats["n_next"]
    assert stats["n_hit"] == stats["n_closest"]  # closed room: no ray escapes
    rad = np.zeros((H, W, 4), f32)
    wgt = np.zeros((H, W), f32)
    s.film_export(rad, wgt)
    assert np.array_equal(wgt, np.full((H, W), float(SPP), f32))
    assert np.isfinite(color).all() and color[..., :3].min() >= 0
    s.film_clear()
    s.render_range(0, npix)
    r0 = np.zeros_like(rad); w0 = np.zeros_like(wgt)
    s.film_export(r0, w0)
    s.film_clear()
    s.render_range(npix, 2 * npix)
    r1 = np.zeros_like(rad); w1 = np.zeros_like(wgt)
    s.film_export(r1, w1)
    assert np.array_equal(w0 + w1, wgt)
    assert np.allclose(r0 + r1, rad, rtol=1e-4, atol=1e-4)
    ref = o.render()
    assert abs(color[..., :3].mean() - ref["color"][..., :3].mean()) / ref["color"][..., :3].mean() < 1e-2
    s.close()
