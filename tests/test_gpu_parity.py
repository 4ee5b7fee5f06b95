"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars: bit-exact for integer / index work (RNG states, pixels, hit flags, primitive ids and the (t, u, v) bits of the
closest hit); fp32 shading values within the tolerances written next to each check (the kernels may contract
a*b+c into FMA and use rsqrt/sincos where the oracle uses IEEE 1/sqrt and sin/cos: SURVEY B.7).
"""
import os

import numpy as np
import pytest
from PIL import Image

from tests import util
from tests.util import GOLDEN, f32

pytestmark = pytest.mark.gpu

KEYS = ("o_med", "d_t1", "pixel", "beta_depth", "path_rad", "rng", "ble")


@pytest.fixture(scope="module")
def pair_small(cornell_small):
    from btrc_b200 import api
    from oracle import front

    s = api.Scene(cornell_small)
    o = front.OracleScene(cornell_small)
    yield s, o
    s.close()


def test_closest_hit_golden_rays(tmp_path, built):
    """lib/bvh's golden image through the CUDA traversal: primitive ids of 777,600 rays, bit-exact vs oracle and vs the PNG."""
    from btrc_b200 import api
    from oracle import front

    path, tris = util.write_fan_cornell_scene(str(tmp_path))
    s = api.Scene(path)
    o = front.OracleScene(path)
    W, H = 1080, 720
    o_med, d_t1 = util.bvh_benchmark_rays(W, H)
    flag, detail = s.closest_hit(o_med, d_t1)
    f0, d0 = o.closest_hit(o_med, d_t1)
    assert np.array_equal(flag, f0)
    assert np.array_equal(detail, d0)  # t bits, primitive id, u bits, v bits
    prim = np.where((flag >> 30) & 1, detail[:, 1].astype(np.int64), -1)
    img = util.normal_image_from_hits(tris.reshape(-1, 9), prim, W, H)
    gold = np.array(Image.open(os.path.join(GOLDEN, "cornell_box_reference.png")))
    assert np.array_equal(img, gold)
    s.close()


def test_generate_parity(pair_small):
    s, o = pair_small
    n = 96 * 96 * 2 + 1000  # crosses a sample-index boundary
    a = s.generate(0, n)
    b = o.generate(0, n)
    assert np.array_equal(a["rng"], b["rng"])          # PCG state + inc: integer, bit-exact
    assert np.array_equal(a["pixel"], b["pixel"])
    assert np.array_equal(a["o_med"], b["o_med"])
    assert np.array_equal(a["beta_depth"].view(np.uint32), b["beta_depth"].view(np.uint32))
    assert np.array_equal(a["path_rad"], b["path_rad"]) and np.array_equal(a["ble"], b["ble"])
    assert np.array_equal(a["d_t1"][:, 3], b["d_t1"][:, 3])
    # direction: rsqrtf (2 ulp) vs 1/sqrt -> 1e-6 absolute on unit vectors
    assert np.abs(a["d_t1"][:, :3] - b["d_t1"][:, :3]).max() < 1e-6
    # a later slice of the index space (sample index 5)
    a = s.generate(96 * 96 * 5 + 17, 500)
    b = o.generate(96 * 96 * 5 + 17, 500)
    assert np.array_equal(a["rng"], b["rng"]) and np.array_equal(a["pixel"], b["pixel"])


def _oracle_bounce_states(o, n, bounces):
    """Advance n camera paths `bounces` times with the oracle; returns the list of (state, flag, detail, shade_out)."""
    st = o.generate(0, n)
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


def test_closest_and_any_hit_on_incoherent_rays(pair_small):
    """Bounce rays (incoherent) and NEE shadow rays produced by the oracle: identical ray batches -> identical hits."""
    s, o = pair_small
    total = 0
    for st, flag, detail, sh in _oracle_bounce_states(o, 96 * 96 * 4, 4):
        f1, d1 = s.closest_hit(st["o_med"], st["d_t1"])
        assert np.array_equal(f1, flag)
        hit = ((flag >> 30) & 1).astype(bool)
        assert np.array_equal(d1[hit], detail[hit])
        has_sh = (sh["status"] & 2) != 0
        so, sd = np.ascontiguousarray(sh["sh_o_med"][has_sh]), np.ascontiguousarray(sh["sh_d_t1"][has_sh])
        assert np.array_equal(s.any_hit(so, sd), o.any_hit(so, sd))
        total += st["o_med"].shape[0] + so.shape[0]
    assert total > 100000


def test_edge_cases_closest_hit(pair_small):
    s, o = pair_small
    # empty batch
    f, d = s.closest_hit(np.zeros((0, 4), f32), np.zeros((0, 4), f32))
    assert f.shape == (0,)
    # ragged batch sizes, rays that miss everything, zero-length intervals, axis-parallel directions
    rng = np.random.default_rng(1)
    for n in (1, 31, 33, 257, 1000):
        o_med = np.zeros((n, 4), f32)
        d_t1 = np.zeros((n, 4), f32)
        o_med[:, :3] = rng.uniform(-0.9, 0.9, (n, 3)) + np.array([0, 1, 0])
        d = rng.standard_normal((n, 3))
        d[::7] = [0, 1, 0]
        d[1::7] = [1, 0, 0]
        d[2::7] = [0, 0, -1]
        d_t1[:, :3] = d
        d_t1[:, 3] = np.finfo(f32).max
        d_t1[3::5, 3] = 0.05       # short rays
        d_t1[4::11, 3] = 0.0       # degenerate interval
        f1, d1 = s.closest_hit(o_med, d_t1)
        f0, d0 = o.closest_hit(o_med, d_t1, brute=True)
        assert np.array_equal(f1, f0)
        hit = ((f0 >> 30) & 1).astype(bool)
        assert np.array_equal(d1[hit], d0[hit])
    # rays starting outside the scene pointing away: all miss
    o_med = np.tile(np.array([[0, 1, 10, 0]], f32), (64, 1))
    d_t1 = np.tile(np.array([[0, 0, 1, np.finfo(f32).max]], f32), (64, 1))
    f1, _ = s.closest_hit(o_med, d_t1)
    assert not f1.any()


def _match(a, b, key_a, key_b):
    ia = np.argsort(key_a, kind="stable")
    ib = np.argsort(key_b, kind="stable")
    assert np.array_equal(key_a[ia], key_b[ib]), "different sets of paths"
    return ia, ib


def test_shade_parity(pair_small):
    """Same states + same hits -> same continuation states, shadow rays and film splats (matched by RNG state)."""
    s, o = pair_small
    for depth, (st, flag, detail, sh) in enumerate(_oracle_bounce_states(o, 96 * 96 * 2, 6)):
        s.film_clear()
        nxt, shd = s.shade(st, flag, detail)
        keep = (sh["status"] & 1) != 0
        ref = {k: v[keep] for k, v in sh["next"].items()}
        assert nxt["rng"].shape[0] == ref["rng"].shape[0], depth
        ia, ib = _match(nxt, ref, nxt["rng"][:, 0], ref["rng"][:, 0])
        for k in ("rng", "pixel"):
            assert np.array_equal(nxt[k][ia], ref[k][ib]), k
        assert np.array_equal(nxt["beta_depth"][ia, 3].view(np.uint32), ref["beta_depth"][ib, 3].view(np.uint32))
        assert np.array_equal(nxt["o_med"][ia, 3].view(np.uint32), ref["o_med"][ib, 3].view(np.uint32))
        # fp32 shading values: relative 2e-4 (FMA contraction, rsqrt, sincos) on throughput-like quantities,
        # 2e-5 absolute on positions / unit directions
        assert np.allclose(nxt["o_med"][ia, :3], ref["o_med"][ib, :3], atol=2e-5)
        # sampled directions: z = sqrt(1 - x^2 - y^2) amplifies the SFU sin/cos error (3.6e-7) by 1 / z towards grazing angles
        dclose = np.isclose(nxt["d_t1"][ia, :3], ref["d_t1"][ib, :3], rtol=0, atol=2e-5).all(axis=1)
        assert dclose.mean() > 0.995, dclose.mean()
        assert np.allclose(nxt["d_t1"][ia, :3], ref["d_t1"][ib, :3], atol=2e-3)
        assert np.allclose(nxt["beta_depth"][ia, :3], ref["beta_depth"][ib, :3], rtol=2e-4, atol=1e-6)
        assert np.allclose(nxt["path_rad"][ia, :3], ref["path_rad"][ib, :3], rtol=2e-4, atol=1e-6)
        # beta_le and the pdf carry cos(theta) = sqrt(1 - sx^2 - sy^2) of the sampled direction, whose relative error grows like
        # 1e-7 / cos^2 towards grazing directions (it cancels in beta = beta_le / pdf above): 2e-4 for 99.5 %, 2 % for all
        close = np.isclose(nxt["ble"][ia], ref["ble"][ib], rtol=2e-4, atol=1e-6).all(axis=1)
        assert close.mean() > 0.995, close.mean()
        assert np.allclose(nxt["ble"][ia], ref["ble"][ib], rtol=2e-2, atol=1e-5)
        # shadow rays: match by (pixel, origin bits) -- one per path at most, paths of one pixel differ in origin
        has_sh = (sh["status"] & 2) != 0
        util.compare_shadow_rays(shd, st, has_sh, sh["sh_o_med"], sh["sh_d_t1"], sh["sh_beta_li"], rtol=5e-4)
        # film: radiance splats of terminated paths
        W = H = 96
        rad = np.zeros((H, W, 4), f32)
        wgt = np.zeros((H, W), f32)
        s.film_export(rad, wgt)
        want = np.zeros((H, W, 3), np.float64)
        term = (sh["status"] & 4) != 0
        np.add.at(want, (st["pixel"][term, 1], st["pixel"][term, 0]), sh["splat"][term].astype(np.float64))
        assert np.allclose(rad[..., :3], want, rtol=1e-3, atol=1e-5)


def test_shadow_stage_parity(pair_small):
    s, o = pair_small
    st, flag, detail, sh = _oracle_bounce_states(o, 96 * 96 * 2, 1)[0]
    has_sh = (sh["status"] & 2) != 0
    shd = dict(pixel=np.ascontiguousarray(st["pixel"][has_sh]), beta_li=np.ascontiguousarray(sh["sh_beta_li"][has_sh]),
               o_med=np.ascontiguousarray(sh["sh_o_med"][has_sh]), d_t1=np.ascontiguousarray(sh["sh_d_t1"][has_sh]))
    s.film_clear()
    s.shadow(shd)
    rad = np.zeros((96, 96, 4), f32)
    wgt = np.zeros((96, 96), f32)
    s.film_export(rad, wgt)
    occ = o.any_hit(shd["o_med"], shd["d_t1"]).astype(bool)
    want = np.zeros((96, 96, 3), np.float64)
    np.add.at(want, (shd["pixel"][~occ, 1], shd["pixel"][~occ, 0]), shd["beta_li"][~occ, :3].astype(np.float64))
    assert np.allclose(rad[..., :3], want, rtol=1e-4, atol=1e-6)


def test_render_matches_oracle_small(pair_small):
    """Whole render at equal spp and sampler seed: event counts and image vs the oracle."""
    s, o = pair_small
    color, _, _, stats = s.render()
    ref = o.render()
    rs = ref["stats"]
    assert stats["n_gen"] == rs["n_gen"] == 96 * 96 * 8
    # paths are draw-for-draw identical except where an fp32 ulp flips a discrete decision (RR / hemisphere tests):
    # event counts agree to 0.1 %
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(stats[k] - rs[k]) <= max(4, 1e-3 * rs[k]), (k, stats[k], rs[k])
    a, b = color[..., :3], ref["color"][..., :3]
    assert np.isfinite(a).all()
    # per-pixel relative MSE at equal spp / seed: < 1e-3 (identical paths up to rare flips; film order is the only noise)
    assert util.rel_mse(a, b) < 1e-3
    assert abs(a.mean() - b.mean()) / b.mean() < 2e-3
    assert np.array_equal(color[..., 3], np.ones((96, 96), f32))


def test_render_c1_full_size_properties(cornell_c1):
    """BASELINE config 1 at full size (512x512, 16 spp): size-independent properties of the render."""
    from btrc_b200 import api

    s = api.Scene(cornell_c1)
    W = H = 512
    color, _, _, stats = s.render()
    assert stats["n_gen"] == W * H * 16
    assert stats["n_closest"] == stats["n_gen"] + stats["n_next"]
    assert stats["n_shadow_miss"] <= stats["n_shadow"] <= stats["n_hit"]
    rad = np.zeros((H, W, 4), f32)
    wgt = np.zeros((H, W), f32)
    s.film_export(rad, wgt)
    assert np.array_equal(wgt, np.full((H, W), 16.0, f32))  # every pixel received exactly spp samples
    assert np.isfinite(color).all() and color[..., :3].min() >= 0
    # resolve == radiance / weight
    assert np.allclose(color[..., :3], rad[..., :3] / 16.0, rtol=1e-6, atol=1e-7)
    # idempotence up to float-atomic ordering: a second render gives the same image within 1e-5 relative
    color2, _, _, stats2 = s.render()
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"):
        assert stats[k] == stats2[k], k
    assert np.allclose(color, color2, rtol=1e-4, atol=1e-5)
    # linearity in the sample index range (the multi-GPU partition): film[0,8) + film[8,16) == film[0,16)
    npix = W * H
    s.film_clear()
    s.render_range(0, 8 * npix)
    r0 = np.zeros_like(rad); w0 = np.zeros_like(wgt)
    s.film_export(r0, w0)
    s.film_clear()
    s.render_range(8 * npix, 16 * npix)
    r1 = np.zeros_like(rad); w1 = np.zeros_like(wgt)
    s.film_export(r1, w1)
    assert np.array_equal(w0 + w1, wgt)
    assert np.allclose(r0 + r1, rad, rtol=1e-4, atol=1e-4)
    # pool-size independence (paths own their RNG streams): a 10x smaller pool renders the same image
    s2 = api.Scene(cornell_c1, state_count=100000)
    color3, _, _, stats3 = s2.render()
    assert stats3["n_closest"] == stats["n_closest"] and stats3["n_shadow"] == stats["n_shadow"]
    assert np.allclose(color, color3, rtol=1e-4, atol=1e-5)
    s2.close()
    s.close()


def test_tile_partition_properties(tmp_path, built):
    """Image-tile partition (BASELINE config 5's wording; partition.tile_path_range): in Tile mode a rank's path range covers
    all samples of a band of pixels, the films of two ranks add up to the whole render, and the image equals the
    Uniform-mode render of the same scene (same (pixel, sample) streams, different order)."""
    import json

    from btrc_b200 import api, partition, scenes

    W, H, spp = 160, 96, 6
    path = scenes.make_cornell(str(tmp_path), width=W, height=H, spp=spp, state_count=30000)
    ref, _, _, stats_ref = api.Scene(path).render()
    root = json.load(open(path))
    root["renderer"]["tile"] = True
    tpath = str(tmp_path / "scene_tile.json")
    json.dump(root, open(tpath, "w"))
    s = api.Scene(tpath)
    color, _, _, stats = s.render()
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert stats[k] == stats_ref[k], k
    assert np.allclose(color, ref, rtol=1e-4, atol=1e-5)
    npix = W * H
    total_r = np.zeros((H, W, 4), f32); total_w = np.zeros((H, W), f32)
    for rank in range(3):  # 3 ranks: bands that end inside a row
        b, e = partition.tile_path_range(rank, 3, spp, npix)
        s.film_clear()
        s.render_range(b, e)
        r = np.zeros((H, W, 4), f32); w = np.zeros((H, W), f32)
        s.film_export(r, w)
        pb, pe = partition.pixel_block(rank, 3, npix)
        mask = np.zeros(npix, bool); mask[pb:pe] = True
        assert np.array_equal(w.reshape(-1), np.where(mask, f32(spp), f32(0)))  # exactly its own pixels, all their samples
        assert not r.reshape(-1, 4)[~mask].any()
        total_r += r; total_w += w
    assert np.array_equal(total_w, np.full((H, W), f32(spp)))
    assert np.allclose(total_r[..., :3] / spp, ref[..., :3], rtol=1e-4, atol=1e-5)
    s.close()


def test_render_vs_oracle_c1_reduced(tmp_path, built):
    """Config 1 geometry at 256x256 / 16 spp against the oracle (full 512x512 costs the CPU ~20 s; same code path)."""
    from btrc_b200 import api, scenes
    from oracle import front

    path = scenes.make_cornell(str(tmp_path), width=256, height=256, spp=16, state_count=300000)
    s = api.Scene(path)
    o = front.OracleScene(path)
    color, _, _, stats = s.render()
    ref = o.render()
    assert util.rel_mse(color[..., :3], ref["color"][..., :3]) < 1e-3
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(stats[k] - ref["stats"][k]) <= 1e-3 * ref["stats"][k], k
    s.close()


def test_gbuffers_and_tile_mode_and_gaussian(tmp_path, built):
    from btrc_b200 import api, scenes
    from oracle import front

    path = scenes.make_cornell(str(tmp_path), width=64, height=48, spp=4, state_count=20000, albedo=True, normal=True, tile=True,
                               filter={"type": "gaussian", "radius": 1.3, "alpha": 1.6})
    s = api.Scene(path)
    o = front.OracleScene(path)
    a = s.generate(0, 64 * 48 * 4)
    b = o.generate(0, 64 * 48 * 4)
    assert np.array_equal(a["rng"], b["rng"]) and np.array_equal(a["pixel"], b["pixel"])  # tile indexing + alias-table draws
    assert np.abs(a["d_t1"][:, :3] - b["d_t1"][:, :3]).max() < 2e-6
    color, albedo, normal, stats = s.render()
    ref = o.render()
    assert util.rel_mse(color[..., :3], ref["color"][..., :3]) < 2e-3
    assert np.allclose(albedo[..., :3], ref["albedo"][..., :3], atol=2e-3)
    assert np.allclose(normal[..., :3], ref["normal"][..., :3], atol=2e-3)
    s.close()


def test_reporter_progress_and_stop(tmp_path, built):
    """Reporter / should_stop through the C ABI (core/reporter.h, core/renderer.h:44-48): monotone percentages from 0 to 100 on
    the rendering thread; a truthy return value stops the render without an image."""
    from btrc_b200 import api, scenes

    path = scenes.make_cornell(str(tmp_path), width=256, height=256, spp=32, state_count=65536)
    s = api.Scene(path)
    seen = []
    s.set_reporter(lambda pct: seen.append(pct) and False)
    color, _, _, stats = s.render()
    assert seen[0] == 0.0 and seen[-1] == 100.0 and len(seen) > 10
    assert all(b >= a for a, b in zip(seen, seen[1:]))
    assert all(b - a >= 1.0 for a, b in zip(seen[1:-2], seen[2:-1]))  # console.cpp:26-33: steps of at least one percent
    ref = color.copy()
    stop_at = []
    s.set_reporter(lambda pct: (stop_at.append(pct) or True) if pct >= 30.0 else False)
    with pytest.raises(api.RenderStopped):
        s.render()
    assert 30.0 <= stop_at[0] < 100.0 and len(stop_at) == 1
    s.set_reporter(None)
    color2, _, _, stats2 = s.render()  # the renderer is reusable after a stop
    assert stats2["n_gen"] == stats["n_gen"] and np.allclose(color2, ref, rtol=1e-4, atol=1e-5)
    s.close()


def test_halton_sampler_parity(tmp_path, built):
    """renderer "sampler": "halton" (builtin/sampler/halton.cpp; the reference picks its sampler at compile time): the kernels'
    Owen-scrambled Halton draws are bit-identical to the oracle's -- sampler states after generate and after a bounce, and the
    whole render."""
    import json

    from btrc_b200 import api, scenes
    from oracle import front

    W, H, spp = 96, 64, 8
    base = scenes.make_cornell(str(tmp_path), width=W, height=H, spp=spp, state_count=20000)
    root = json.load(open(base))
    root["renderer"]["sampler"] = "halton"
    path = str(tmp_path / "halton.json")
    json.dump(root, open(path, "w"))
    s = api.Scene(path)
    o = front.OracleScene(path)
    n = W * H * 2 + 777
    a, b = s.generate(0, n), o.generate(0, n)
    assert np.array_equal(a["rng"], b["rng"]) and np.array_equal(a["pixel"], b["pixel"])      # (halton index, dimension << 1)
    assert (a["rng"][:, 1] == 5 << 1).all()
    assert np.abs(a["d_t1"][:, :3] - b["d_t1"][:, :3]).max() < 1e-6                          # same film offsets => same rays
    # one shade step on identical states: the continuation states carry identical sampler states and matching throughput
    flag, detail = o.closest_hit(b["o_med"], b["d_t1"])
    ref = o.shade(b, flag, detail)
    nxt, shd = s.shade(b, flag, detail)
    keep = (ref["status"] & 1) != 0
    assert abs(nxt["rng"].shape[0] - int(keep.sum())) <= 2
    ka = {(int(p[0]), int(p[1]), int(r[0])): i for i, (p, r) in enumerate(zip(nxt["pixel"], nxt["rng"]))}
    hits = 0
    for j in np.nonzero(keep)[0][:5000]:
        key = (int(ref["next"]["pixel"][j][0]), int(ref["next"]["pixel"][j][1]), int(ref["next"]["rng"][j][0]))
        i = ka.get(key)
        if i is None:
            continue
        hits += 1
        assert nxt["rng"][i][1] == ref["next"]["rng"][j][1]                                   # same number of draws
        assert np.allclose(nxt["beta_depth"][i, :3], ref["next"]["beta_depth"][j, :3], rtol=2e-3, atol=1e-5)
    assert hits > 4000
    color, _, _, stats = s.render()
    r = o.render()
    assert stats["n_gen"] == r["stats"]["n_gen"]
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(stats[k] - r["stats"][k]) <= 2e-3 * r["stats"][k] + 4, k
    assert util.rel_mse(color[..., :3], r["color"][..., :3]) < 2e-3
    s.close()


def test_device_lbvh_builder(tmp_path, built):
    """renderer "bvh_builder": "lbvh" (device/lbvh.cu: Morton codes, radix sort, Karras radix tree, bottom-up fit) -- the hits do not
    depend on the hierarchy: the 777,600 golden rays and a whole render give exactly what the host SAH tree gives."""
    import json

    from btrc_b200 import api, scenes
    from oracle import front

    path, tris = util.write_fan_cornell_scene(str(tmp_path))[:2]
    root = json.load(open(path))
    root["renderer"]["bvh_builder"] = "lbvh"
    lpath = str(tmp_path / "fan_lbvh.json")
    json.dump(root, open(lpath, "w"))
    s = api.Scene(lpath)
    o = front.OracleScene(path)
    o_med, d_t1 = util.bvh_benchmark_rays()
    f0, d0 = o.closest_hit(o_med, d_t1)
    f1, d1 = s.closest_hit(o_med, d_t1)
    assert np.array_equal(f0, f1) and np.array_equal(d0, d1)
    s.close()
    # a scene with many triangles: materials test (15 spheres), render with both builders
    base = scenes.make_material_test(str(tmp_path / "m"), width=160, height=90, spp=4, rings=40, segments=40, tex_res=64, state_count=40000)
    root = json.load(open(base))
    root["renderer"]["bvh_builder"] = "lbvh"
    lbase = os.path.join(os.path.dirname(base), "scene_lbvh.json")
    json.dump(root, open(lbase, "w"))
    a = api.Scene(base)
    b = api.Scene(lbase)
    ca, _, _, sa = a.render()
    cb, _, _, sb = b.render()
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"):
        assert sa[k] == sb[k], k
    assert np.allclose(ca, cb, rtol=1e-4, atol=1e-5)
    assert b.info.bvh_build_seconds > 0
    st = front.OracleScene(base).generate(0, 160 * 90)
    fa, da = a.closest_hit(st["o_med"], st["d_t1"])
    fb, db = b.closest_hit(st["o_med"], st["d_t1"])
    assert np.array_equal(fa, fb) and np.array_equal(da, db)
    a.close(); b.close()


@pytest.mark.parametrize("mode", ["1", "2"])
def test_binned_ray_order_is_transparent(tmp_path, built, mode, monkeypatch):
    """The binned visiting order of surviving paths and shadow rays (types.h: RaySort; on by default only when the BVH exceeds the
    L2) changes no result: same event counts, same image up to float-atomic ordering, same traversal form or the other."""
    from btrc_b200 import api, scenes

    path = scenes.make_material_test(str(tmp_path), width=160, height=90, spp=8, rings=24, segments=32, tex_res=64, state_count=30000)
    monkeypatch.setenv("BTRC_SORT_RAYS", "0")
    a = api.Scene(path)
    ca, _, _, sa = a.render()
    monkeypatch.setenv("BTRC_SORT_RAYS", mode)
    monkeypatch.setenv("BTRC_TRAVERSE", "ifif" if mode == "2" else "ww")
    b = api.Scene(path)
    cb, _, _, sb = b.render()
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"):
        assert sa[k] == sb[k], k
    assert np.allclose(ca, cb, rtol=1e-4, atol=1e-5)
    a.close(); b.close()
