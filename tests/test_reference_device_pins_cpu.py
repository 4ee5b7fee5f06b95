"""Pins of the ORACLE against the REFERENCE'S OWN device code, executed on the host.

oracle/_ref/libbtrcrefcpu.so is Btrc's wavefront path tracer compiled from /root/reference where it lies: its JSON factory,
scene commit, `GeneratePipeline` / `MediumPipeline` / `ShadePipeline::record_device_code` with every material, medium, light,
filter and the camera they call, printed as C++ by CUJ's own LLVM-free `CPPCodeGenerator` and compiled with g++ (IEEE fp32, no
contraction: the oracle's contract).  oracle/refcpu/refcpu.cpp lists the three stand-ins (OptiX -> the reference's vendored
lib/bvh; the texture unit; the OptiX getters of the shadow miss program).

What is compared, on identical batches, for every scene type the path has: generate (states, rays, RNG), closest hit, then
three bounces of [medium ->] shade: statuses, next states, shadow rays, film splats, sampler states.  The bar is BIT-EXACT for
every integer AND every fp32 value -- the oracle restates the reference's arithmetic operation for operation.  Two documented
differences are normalised before comparing, nothing else:
  * medium id of "no medium": the reference assigns one past its table (core/scene.cpp:160); oracle and product
    use the volume-primitive id just below it (`ReferenceScene.to_oracle_ids`);
  * `AliasTable::initialize` sums with std::reduce, whose association order is the standard library's: the reference built here
    has libstdc++'s (groups of four), the oracle defaults to MSVC's sequential order; `orc_set_reduce_order(1)` selects the
    former for these comparisons (tests/test_reference_host_pins_cpu.py pins both).
"""
import numpy as np
import pytest

from tests.util import f32

refcpu = pytest.importorskip("oracle.refcpu")
pytestmark = pytest.mark.skipif(not refcpu.available(), reason="oracle/_ref/libbtrcrefcpu.so not built (needs /root/reference: make -C oracle refcpu)")

STATE_KEYS = ("o_med", "d_t1", "pixel", "beta_depth", "path_rad", "rng", "ble")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(a.shape[0], -1) if a.ndim > 1 else a.view(np.uint8).reshape(a.shape[0], -1)


def _assert_bits(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, what
    if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
        bad = np.nonzero((a != b).reshape(a.shape[0], -1).any(axis=1))[0]
        raise AssertionError("%s: %d of %d rows differ, first %d: oracle %r reference %r" % (what, bad.size, a.shape[0], bad[0], a[bad[0]], b[bad[0]]))


def _make(kind, d):
    from btrc_b200 import scenes

    if kind == "cornell":
        return scenes.make_cornell(d, width=64, height=64, spp=4, state_count=20000)
    if kind == "materials":  # disney (textured, normal-mapped, clearcoat / sheen / anisotropy / transmission sweeps), metal, glass, mirror, diffuse lights, sky
        return scenes.make_material_test(d, width=96, height=54, spp=4, rings=12, segments=16, tex_res=64, state_count=60000)
    if kind == "medium":     # heterogeneous grid medium inside a transform, HG phase function, invisible boundary
        return scenes.make_medium_scene(d, width=64, height=64, spp=4, grid_res=24, state_count=50000)
    if kind == "homogeneous":
        return scenes.make_medium_scene(d, width=64, height=64, spp=4, grid_res=24, state_count=50000, heterogeneous=False)
    if kind == "ibl":        # image-based light: the LUT comes from the reference's own preprocess kernel (env_sampler.cpp:22-122)
        return scenes.make_ibl_scene(d, width=96, height=54, spp=4, rings=10, segments=14, env_res=(128, 64), lut=(32, 16), state_count=40000)
    if kind == "volumes":    # two overlapping volume primitives
        return scenes.make_volume_scene(d, width=64, height=48, spp=4, grid_res=16, state_count=40000)
    if kind == "instanced":  # BASELINE config 4 in small: one mesh under 10 rigid transforms + scales, diffuse / metal, 6 quad lights, depth 16
        return scenes.make_instanced_scene(d, width=96, height=54, spp=4, rings=13, segments=24, n_instances=10, n_lights=6, state_count=60000)
    if kind == "extras":
        return _make_extras(d)
    if kind == "medium_bytes":
        return _make_medium_bytes(d)
    raise ValueError(kind)


def _make_medium_bytes(d):
    """The medium scene with a "byte" text grid: the reference extracts one CHARACTER per component and uploads SNorm8
    (utils/cuda/array.cpp:331-355), i.e. densities char / 127; scaled by a `mul` node to a useful optical depth, plus a
    three-channel byte grid as albedo (fourth component 255 = -1 / 127)."""
    import json
    import os

    from btrc_b200 import scenes

    path = scenes.make_medium_scene(d, width=64, height=64, spp=4, grid_res=24, state_count=50000)
    rng = np.random.default_rng(23)
    res = 12
    def grid(channels):
        chars = rng.integers(40, 127, res * res * res * channels)          # printable, non-blank ASCII
        body = "".join(chr(c) for c in chars)
        lines = [body[i:i + 97] for i in range(0, len(body), 97)]          # blanks and newlines between values are skipped
        return "byte %d %d %d %d\n" % (channels, res, res, res) + " \n".join(lines) + "\n"
    open(os.path.join(d, "density_bytes.txt"), "w").write(grid(1))
    open(os.path.join(d, "albedo_bytes.txt"), "w").write(grid(3))
    root = json.load(open(path))
    for e in root["scene"]["entities"]:
        m = e.get("inner_medium")
        if m:
            m["medium"]["sigma_t"] = {"type": "mul", "lhs": {"type": "array", "text": "$env{scene_directory}/density_bytes.txt"}, "rhs": 6.0}
            m["medium"]["albedo"] = {"type": "array", "text": "$env{scene_directory}/albedo_bytes.txt", "filter": "point"}
    json.dump(root, open(path, "w"), indent=1)
    return path


def _make_extras(d):
    """What the other scenes leave out: `black` and `invisible` surfaces, a texture graph (mul / add / transform), mirror and
    border address modes with a point filter, g-buffer outputs, the box filter with tile-mode indexing.  (One Transform2D op,
    not an array: the reference's Mat3 product reads a fourth row and column that do not exist, mat3.cpp:108-123.)"""
    import json
    import os

    from btrc_b200 import scenes

    base, rough, nmap = scenes.write_test_textures(d, 32)
    v, f, uv, n = scenes.uv_sphere(8, 12)
    scenes.write_obj(os.path.join(d, "sphere.obj"), v, f, uv, n)
    gv, gf, guv = scenes.quad((-6, 0, 6), (6, 0, 6), (6, 0, -6), (-6, 0, -6))
    scenes.write_obj(os.path.join(d, "ground.obj"), gv, gf, [(u * 3 - 1, w * 3 - 1) for u, w in guv])
    lv, lf, luv = scenes.quad((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1))
    scenes.write_obj(os.path.join(d, "lightquad.obj"), lv, lf, luv)
    tex = lambda name, **kw: dict({"type": "array", "filename": "$env{scene_directory}/" + name}, **kw)
    graph = {"type": "mul", "lhs": {"type": "add", "lhs": tex(base, address_mode_u="mirror", address_mode_v="mirror"), "rhs": [0.05, 0.1, 0.02]},
             "rhs": {"type": "transform", "transform": {"rotate": {"angle": {"deg": 30}}}, "tex": tex(rough, address_mode_u="border", address_mode_v="wrap", filter="point", border=[0.5])}}
    geo = lambda name: {"type": "triangle_mesh", "filename": "$env{scene_directory}/%s.obj" % name}
    ents = [{"geometry": geo("ground"), "material": {"type": "diffuse", "color": graph, "normal": tex(nmap, address_mode_u="mirror", address_mode_v="border")}}]
    mats = [{"type": "black"}, {"type": "invisible"}, {"type": "diffuse", "color": graph, "shadow_terminator_term": False},
            {"type": "metal", "color": graph, "roughness": tex(rough, address_mode_u="wrap", address_mode_v="mirror"), "anisotropic": 0.3}]
    for i, m in enumerate(mats):
        ents.append({"geometry": geo("sphere"), "material": m, "local_to_world": [{"translate": [2.4 * (i - 1.5), 1.0, 0.0]}]})
    ents.append({"geometry": geo("lightquad"), "material": {"type": "diffuse", "color": [0, 0, 0]},
                 "local_to_world": [{"scale": [2, 1, 2]}, {"translate": [0, 5, 0]}], "light": {"type": "mesh_light", "intensity": [30, 30, 28]}})
    root = {"width": 80, "height": 48,
            "camera": {"type": "pinhole", "eye": [0, 3.5, 8], "dst": [0, 0.8, 0], "up": [0, 1, 0], "fov_y": {"deg": 40}},
            "renderer": {"type": "wfpt", "spp": 4, "state_count": 40000, "min_depth": 3, "max_depth": 6, "albedo": True, "normal": True, "tile": True,
                         "filter": {"type": "box"}},
            "scene": {"entities": ents, "envir_light": {"type": "gradient_sky", "upper": [0.3, 0.4, 0.6], "lower": [0.1, 0.1, 0.1], "up": [0, 1, 0]}}}
    path = os.path.join(d, "scene.json")
    with open(path, "w") as fjson:
        json.dump(root, fjson, indent=1)
    return path


KINDS = ["cornell", "materials", "medium", "homogeneous", "ibl", "volumes", "extras", "medium_bytes", "instanced"]


@pytest.fixture(scope="module", params=KINDS)
def pair(request, tmp_path_factory, built):
    from oracle import front

    front.lib().orc_set_reduce_order(1)  # the reference as compiled here sums with libstdc++'s std::reduce
    d = str(tmp_path_factory.mktemp("refpin_" + request.param))
    path = _make(request.param, d)
    R = refcpu.ReferenceScene(path)
    O = front.OracleScene(path)
    yield request.param, O, R
    R.close()
    front.lib().orc_set_reduce_order(0)


def _compare_stage(o, r, what, gbuffer=False):
    assert np.array_equal(o["status"] & 3, r["status"] & 3), what + ": next / shadow emission decisions"
    nxt, sh = (o["status"] & 1) != 0, (o["status"] & 2) != 0
    for k in STATE_KEYS:
        _assert_bits(o["next"][k][nxt], r["next"][k][nxt], "%s next.%s" % (what, k))
    for k in ("sh_o_med", "sh_d_t1", "sh_beta_li"):
        _assert_bits(o[k][sh], r[k][sh], "%s %s" % (what, k))
    # the reference kernel leaves no trace of a zero splat, so compare values: where the oracle splats, the film must have
    # received exactly that; elsewhere nothing (non-finite splats are dropped by both, film.cpp:131-165)
    rad = (o["status"] & 4) != 0
    want = np.where(rad[:, None], o["splat"], 0).astype(f32)
    want[~np.isfinite(want).all(axis=1)] = 0
    _assert_bits(want, r["splat"], what + " radiance splat")
    if gbuffer:
        g = (o["status"] & 8) != 0
        _assert_bits(np.where(g[:, None], o["albedo"], 0).astype(f32), r["albedo"], what + " g-buffer albedo")
        _assert_bits(np.where(g[:, None], o["normal"], 0).astype(f32), r["normal"], what + " g-buffer normal")


def test_geometry_and_tables_match(pair):
    """World-space triangles (= primitive ids) and the six float4 shading tables per triangle (core/geometry.h:10-19)."""
    kind, O, R = pair
    ov, oip = O.triangles()
    rv, rip = R.triangles()
    _assert_bits(ov, rv, kind + " world triangles")
    assert np.array_equal(oip, rip)
    assert O.vol_prim_medium_id() == R.vol_prim_medium_id() and O.has_medium() == R.has_medium()


def test_generate_matches_reference(pair):
    """generate.cpp:39-86 with the reference's own filter / camera / sampler code: every byte of every state."""
    kind, O, R = pair
    n = R.width * R.height * 2
    first = R.width * R.height - 7  # straddles a sample-index boundary
    so, sr = O.generate(first, n), R.generate(first, n)
    for k in STATE_KEYS:
        _assert_bits(so[k], sr[k], "%s generate.%s" % (kind, k))


def _stages_of_batch(kind, O, R, first, n, seen, label):
    """generate [first, first + n) and follow it for three bounces of closest hit, [medium,] shade; returns the states compared."""
    st = O.generate(first, n)
    total = 0
    for bounce in range(3):
        what = "%s %s bounce %d" % (kind, label, bounce)
        fo, do = O.closest_hit(st["o_med"], st["d_t1"])
        fr, dr = R.closest_hit(st["o_med"], st["d_t1"])
        assert np.array_equal(fo, fr) and np.array_equal(do, dr), what + " closest hit"
        total += st["o_med"].shape[0]
        mo = None
        if R.has_medium():
            so2, fo2, mo = O.medium(st, fo, do)
            sr2, fr2, mr = R.medium(st, fo, do)
            assert np.array_equal(fo2, fr2), what + " medium: scattered flags"
            R.to_oracle_ids(sr2["o_med"])
            for k in STATE_KEYS:
                _assert_bits(so2[k], sr2[k], "%s medium in-place %s" % (what, k))
            R.to_oracle_ids(mr["next"]["o_med"]); R.to_oracle_ids(mr["sh_o_med"])
            _compare_stage(mo, mr, what + " medium")
            st2, fl = so2, fo2
        else:
            st2, fl = st, fo
        oo, rr = O.shade(st2, fl, do), R.shade(st2, fl, do)
        R.to_oracle_ids(rr["next"]["o_med"]); R.to_oracle_ids(rr["sh_o_med"])
        _compare_stage(oo, rr, what + " shade", gbuffer=(kind == "extras"))
        seen += np.bincount(oo["status"], minlength=16)[:16]
        keep = (oo["status"] & 1) != 0
        parts = [{k: v[keep] for k, v in oo["next"].items()}]
        if mo is not None:
            km = (mo["status"] & 1) != 0
            parts.append({k: v[km] for k, v in mo["next"].items()})
        st = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in parts[0]}
        assert st["o_med"].shape[0] > (100 if bounce == 0 else 16), what + ": too few surviving paths to mean anything"
    return total


def test_stages_match_reference_bit_for_bit(pair):
    """Closest hit, medium and shade over three bounces: oracle == reference's own device code, bit for bit -- on at least 250,000
    states per scene type (batch after batch of camera paths, sample indices running past the scene's spp: the sampler does not care;
    the tile-mode scene stops at its own spp)."""
    kind, O, R = pair
    batch = R.width * R.height * 2
    seen = np.zeros(16, np.int64)
    total, first, k = 0, 0, 0
    # "extras" renders in tile mode (index = pixel * spp + sample): indices past W * H * spp name no pixel, so it stops at its 4 spp
    limit = R.width * R.height * 4 if kind == "extras" else None
    while total < 250000 and (limit is None or first + batch <= limit):
        total += _stages_of_batch(kind, O, R, first, batch, seen, "batch %d" % k)
        first += batch
        k += 1
    assert total >= (250000 if limit is None else 20000)
    assert seen[1:].sum() > batch, "stages exercised"


def test_shadow_stage_matches_reference(pair):
    """shadow.cpp:24-76: occlusion, transmittance along unoccluded shadow rays, film splat.  Heterogeneous media draw their
    ratio-tracking samples from per-slot generators in the reference (soa_buffer.cpp:86-96) and from a ray-keyed stream in
    oracle and product (DESIGN: documented deviation), so those are compared in the mean; everything else bit for bit."""
    kind, O, R = pair
    n = R.width * R.height
    st = O.generate(n * R.spp // 2, n)  # the middle of the index range: tile-mode indexing walks pixel by pixel
    fo, do = O.closest_hit(st["o_med"], st["d_t1"])
    if R.has_medium():
        st, fo, _ = O.medium(st, fo, do)
    oo = O.shade(st, fo, do)
    sh = (oo["status"] & 2) != 0
    assert sh.sum() > 200
    so, sd, sb, px = oo["sh_o_med"][sh], oo["sh_d_t1"][sh], oo["sh_beta_li"][sh], st["pixel"][sh]
    occ_o = O.any_hit(so, sd)
    tr = O.shadow_tr(so, sd)
    occ_r, contrib = R.shadow(so, sd, sb, px)
    assert np.array_equal(occ_o != 0, occ_r != 0), kind + " occlusion"
    want = (sb[:, :3] * tr).astype(f32)
    want[occ_o != 0] = 0
    want[~np.isfinite(want).all(axis=1)] = 0
    if kind in ("medium", "volumes", "medium_bytes"):
        m = occ_o == 0
        assert abs(contrib[m].mean() - want[m].mean()) <= 0.05 * want[m].mean() + 1e-6
        # rays that meet no density get tr = 1 from any stream: most of them, and those must agree exactly
        same = (want[m].view(np.uint32) == contrib[m].view(np.uint32)).all(axis=1)
        assert same.mean() > 0.5, same.mean()
    else:
        _assert_bits(want, contrib, kind + " shadow contribution")


def test_render_matches_reference(pair):
    """Whole render, wavefront.cpp:205-325 against the oracle's loop: identical event counts (so identical decisions for every
    path of the image) and an image equal up to the order of the film's float additions."""
    kind, O, R = pair
    ro, rr = O.render(), R.render()
    a, b = ro["color"][..., :3], rr["color"][..., :3]
    if kind == "volumes":
        # rays without a medium carry the reference's out-of-table id and skip the volume primitives there (its access_medium
        # falls into `unreachable`); oracle and product give them the volume-primitive medium.  Same light transport otherwise.
        assert abs(a.mean() - b.mean()) < 0.03 * a.mean()
        return
    for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        if kind in ("medium", "medium_bytes") and k != "n_gen":
            assert abs(ro["stats"][k] - rr["stats"][k]) <= 0.01 * ro["stats"][k], k  # shadow-ray Tr streams differ (see above) only in the film
        else:
            assert ro["stats"][k] == rr["stats"][k], (kind, k, ro["stats"][k], rr["stats"][k])
    if kind in ("medium", "medium_bytes"):
        assert ((a - b) ** 2).mean() / (a.mean() ** 2) < 5e-2
    else:
        assert np.abs(a - b).max() <= 1e-5 * max(1.0, float(a.max())), np.abs(a - b).max()
