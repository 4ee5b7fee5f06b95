"""CPU tests of the oracle's material / texture / medium restatement (no GPU).

The reference holds no fixtures for these parts (parity unpinned by the reference), so the oracle is pinned by
(a) internal consistency laws every BSDF of the reference satisfies by construction (a sample's value and pdf are the
eval / pdf of its direction; pdfs integrate to <= 1; reciprocity of the symmetric lobes), (b) closed forms (homogeneous
transmittance, constant-density ratio tracking) and (c) frozen known-answer vectors (tests/golden/bsdf_kat.json, written by
tests/golden/make_golden_bsdf.py from the oracle itself: regression pins).
"""
import json
import os

import numpy as np
import pytest

from tests.util import GOLDEN, f32

MATERIALS = {
    "diffuse": {"type": "diffuse", "color": [0.7, 0.5, 0.3]},
    "metal": {"type": "metal", "color": [0.95, 0.64, 0.54], "roughness": 0.35, "anisotropic": 0.4},
    "disney_plastic": {"type": "disney", "color": [0.8, 0.3, 0.2], "metallic": 0.0, "roughness": 0.45, "clearcoat": 0.7, "clearcoat_gloss": 0.8,
                       "sheen": 0.5, "sheen_tint": 0.5, "specular_tint": 0.3},
    "disney_metal": {"type": "disney", "color": [0.9, 0.8, 0.4], "metallic": 1.0, "roughness": 0.3, "anisotropic": 0.6},
    "disney_glass": {"type": "disney", "color": [0.9, 0.9, 0.95], "metallic": 0.0, "roughness": 0.3, "transmission": 0.9, "ior": 1.45,
                     "transmission_roughness": 0.25},
    "glass": {"type": "glass", "color": [0.95, 0.97, 1.0], "ior": 1.5},
    "mirror": {"type": "mirror", "color": [0.9, 0.85, 0.7]},
    "invisible": {"type": "invisible"},
    "black": {"type": "black"},
}


def material_scene(tmp_dir, extra_entities=()):
    """One quad entity per test material (the geometry is irrelevant: bsdf_probe evaluates materials at a synthetic hit)."""
    from btrc_b200 import scenes

    os.makedirs(tmp_dir, exist_ok=True)
    v, f, uv = scenes.quad((-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1))
    scenes.write_obj(os.path.join(tmp_dir, "q.obj"), v, f, uv)
    ents = [{"geometry": {"type": "triangle_mesh", "filename": "$env{scene_directory}/q.obj"}, "material": m,
             "local_to_world": {"translate": [3.0 * i, 0, 0]}} for i, m in enumerate(MATERIALS.values())]
    ents += list(extra_entities)
    root = {"width": 16, "height": 16, "camera": {"type": "pinhole", "eye": [0, 3, 5], "dst": [0, 0, 0], "up": [0, 1, 0], "fov_y": {"deg": 40}},
            "renderer": {"type": "wfpt", "spp": 1}, "scene": {"entities": ents}}
    path = os.path.join(tmp_dir, "scene.json")
    json.dump(root, open(path, "w"))
    return path


@pytest.fixture(scope="module")
def mat_oracle(tmp_path_factory, built):
    from oracle import front

    return front.OracleScene(material_scene(str(tmp_path_factory.mktemp("mats"))))


def _dirs(rng, n, upper=None):
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if upper is True:
        d[:, 2] = np.abs(d[:, 2])
    return d.astype(f32)


@pytest.mark.parametrize("name", ["diffuse", "metal", "disney_plastic", "disney_metal", "disney_glass"])
def test_sample_is_consistent_with_eval_and_pdf(mat_oracle, name):
    """BSDFAggregate::sample returns eval(dir) and pdf(dir) of the direction it picked (aggregate.cpp:84-117)."""
    mid = list(MATERIALS).index(name)
    rng = np.random.default_rng(5)
    n = 4000
    wo = _dirs(rng, n, upper=None if name == "disney_glass" else True)
    sam = rng.random((n, 3)).astype(f32)
    interp_z = np.array([0.1, -0.05, 1.0], f32)
    r = mat_oracle.bsdf_probe(mid, interp_z, [0.3, 0.6], wo, wo, sam)
    ok = r["sample"][:, 6] > 0
    assert ok.mean() > (0.5 if name == "disney_glass" else 0.8)
    dirs = np.ascontiguousarray(r["sample"][ok, 3:6])
    r2 = mat_oracle.bsdf_probe(mid, interp_z, [0.3, 0.6], np.ascontiguousarray(wo[ok]), dirs, np.ascontiguousarray(sam[ok]))
    # the aggregate evaluates the closure in local space on the sampled local direction; re-evaluating from the global
    # direction differs by the round trip through the frame (1e-6 relative, more for very peaked lobes)
    pclose = np.isclose(r2["pdf"], r["sample"][ok, 6], rtol=2e-2, atol=1e-6)
    close = np.isclose(r2["eval"], r["sample"][ok, 0:3], rtol=2e-2, atol=1e-6).all(axis=1)
    assert pclose.mean() > 0.98 and close.mean() > 0.98, (pclose.mean(), close.mean())
    # the well-conditioned quantity is the path weight f / pdf
    w1 = r["sample"][ok, 0:3] / r["sample"][ok, 6:7]
    w2 = r2["eval"] / np.maximum(r2["pdf"][:, None], 1e-20)
    assert np.isclose(w1, w2, rtol=2e-2, atol=1e-5).all(axis=1).mean() > 0.995


@pytest.mark.parametrize("name", ["diffuse", "metal", "disney_plastic", "disney_metal", "disney_glass"])
def test_pdf_integrates_to_at_most_one(mat_oracle, name):
    mid = list(MATERIALS).index(name)
    rng = np.random.default_rng(11)
    n = 400000
    wi = _dirs(rng, n)
    for wo in ([0.0, 0.0, 1.0], [0.6, 0.0, 0.8], [0.5, 0.5, 0.7071]):
        wo_arr = np.tile(np.array(wo, f32), (n, 1))
        r = mat_oracle.bsdf_probe(mid, [0, 0, 1], [0.5, 0.5], wo_arr, wi, np.zeros((n, 3), f32))
        integral = float(r["pdf"].astype(np.float64).mean() * 4 * np.pi)
        # the transmission lobe's pdf (disney.cpp:369-404) omits cos(theta_h): not exactly normalised in the reference either
        assert 0.5 < integral < (1.12 if name == "disney_glass" else 1.03), (name, wo, integral)


@pytest.mark.parametrize("name", ["diffuse", "metal", "disney_plastic", "disney_metal"])
def test_energy_and_reciprocity(mat_oracle, name):
    mid = list(MATERIALS).index(name)
    rng = np.random.default_rng(13)
    n = 200000
    wo = np.tile(np.array([0.3, 0.2, 0.9327], f32), (n, 1))
    sam = rng.random((n, 3)).astype(f32)
    r = mat_oracle.bsdf_probe(mid, [0, 0, 1], [0.5, 0.5], wo, wo, sam)
    ok = r["sample"][:, 6] > 0
    w = r["sample"][ok, 0:3] * np.abs(r["sample"][ok, 5:6]) / r["sample"][ok, 6:7]
    albedo = w.astype(np.float64).sum(axis=0) / n
    assert (albedo < 1.05).all() and (albedo > 0.02).all(), albedo  # no energy gain (the Disney clearcoat may add a little)
    # reciprocity f(wi, wo) == f(wo, wi) for the reflection lobes (shading normal == geometric normal here)
    a, b = _dirs(rng, 2000, True), _dirs(rng, 2000, True)
    f_ab = mat_oracle.bsdf_probe(mid, [0, 0, 1], [0.5, 0.5], a, b, np.zeros((2000, 3), f32))["eval"]
    f_ba = mat_oracle.bsdf_probe(mid, [0, 0, 1], [0.5, 0.5], b, a, np.zeros((2000, 3), f32))["eval"]
    assert np.allclose(f_ab, f_ba, rtol=2e-3, atol=1e-5)


def test_delta_materials(mat_oracle):
    rng = np.random.default_rng(17)
    n = 2000
    wo = _dirs(rng, n, True)
    sam = rng.random((n, 3)).astype(f32)
    names = list(MATERIALS)
    # mirror: perfect reflection about the normal, delta, pdf 1; eval / pdf are zero
    r = mat_oracle.bsdf_probe(names.index("mirror"), [0, 0, 1], [0, 0], wo, wo, sam)
    assert np.allclose(r["sample"][:, 3:6], wo * np.array([-1, -1, 1], f32), atol=1e-6)
    assert (r["sample"][:, 7] == 1).all() and np.allclose(r["sample"][:, 6], 1) and not r["eval"].any() and not r["pdf"].any()
    # invisible: straight through with bsdf = 1 / |cos|, so that beta * |cos| * bsdf / pdf == beta
    r = mat_oracle.bsdf_probe(names.index("invisible"), [0, 0, 1], [0, 0], wo, wo, sam)
    assert np.allclose(r["sample"][:, 3:6], -wo, atol=1e-7)
    assert np.allclose(r["sample"][:, 0] * np.abs(wo[:, 2]), 1, rtol=1e-5)
    # glass: reflect with probability F, refract otherwise (Snell), both delta
    r = mat_oracle.bsdf_probe(names.index("glass"), [0, 0, 1], [0, 0], wo, wo, sam)
    refl = r["sample"][:, 5] > 0
    assert refl.any() and (~refl).any() and (r["sample"][:, 7] == 1).all()
    sin_i = np.sqrt(1 - wo[~refl, 2] ** 2)
    sin_t = np.sqrt(np.maximum(0, 1 - r["sample"][~refl, 5] ** 2))
    assert np.allclose(sin_i, 1.5 * sin_t, atol=2e-5)
    # black: nothing
    r = mat_oracle.bsdf_probe(names.index("black"), [0, 0, 1], [0, 0], wo, wo, sam)
    assert not r["sample"][:, :7].any() and not r["albedo"].any()


def test_bsdf_known_answer_vectors(mat_oracle):
    """Frozen oracle outputs (regression pin): tests/golden/bsdf_kat.json."""
    kat = json.load(open(os.path.join(GOLDEN, "bsdf_kat.json")))
    for name, rec in kat["materials"].items():
        mid = list(MATERIALS).index(name)
        wo, wi, sam = (np.array(kat[k], f32) for k in ("wo", "wi", "sam"))
        r = mat_oracle.bsdf_probe(mid, kat["interp_z"], kat["uv"], wo, wi, sam)
        # libm sin/cos/pow may differ in the last bit between hosts: 1e-5 relative
        assert np.allclose(r["sample"], np.array(rec["sample"], f32), rtol=1e-5, atol=1e-7), name
        assert np.allclose(r["eval"], np.array(rec["eval"], f32), rtol=1e-5, atol=1e-7), name
        assert np.allclose(r["pdf"], np.array(rec["pdf"], f32), rtol=1e-5, atol=1e-7), name


def test_texture_filtering_rule(tmp_path, built):
    """The published CUDA filtering rule on a tiny image: texel centres reproduce texels, address modes, sRGB decode."""
    from PIL import Image

    from oracle import front

    img = np.array([[[0, 0, 0], [255, 0, 0]], [[0, 255, 0], [0, 0, 255]]], np.uint8)
    Image.fromarray(img, "RGB").save(tmp_path / "t.png")
    o = front.OracleScene()
    o.scene_dir = str(tmp_path)

    def tex(**kw):
        node = front.Node(front._json_to_tree(dict({"type": "array", "filename": "$env{scene_directory}/t.png"}, **kw)))
        return o.add_texture(node)

    t_clamp, t_wrap, t_point, t_srgb = tex(), tex(address_mode_u="wrap", address_mode_v="wrap"), tex(filter="point"), tex(srgb_to_linear=True)
    centres = np.array([[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.75, 0.75]], f32)
    want = img.reshape(4, 3).astype(f32) / 255
    assert np.allclose(o.eval_texture2d(t_clamp, centres), want, atol=1e-6)
    assert np.allclose(o.eval_texture2d(t_point, centres + 0.2), want, atol=1e-6)
    mid = o.eval_texture2d(t_clamp, np.array([[0.5, 0.5]], f32))[0]
    assert np.allclose(mid, want.mean(axis=0), atol=1e-6)
    # clamp: beyond the edge stays at the edge texel; wrap: u + 1 is the same point, and u = 0 blends with the far texel
    assert np.allclose(o.eval_texture2d(t_clamp, np.array([[-3.0, 0.25]], f32))[0], want[0], atol=1e-6)
    assert np.allclose(o.eval_texture2d(t_wrap, centres + 1), want, atol=1e-6)
    assert np.allclose(o.eval_texture2d(t_wrap, np.array([[0.0, 0.25]], f32))[0], 0.5 * (want[0] + want[1]), atol=1e-6)
    lin = o.eval_texture2d(t_srgb, centres)
    assert np.allclose(lin, want, atol=1e-6)  # 0 and 1 are fixed points of the sRGB curve
    # operator and transform textures (texture2d/math.cpp, transform.cpp)
    add = o.add_texture(front.Node(front._json_to_tree({"type": "add", "lhs": [0.25, 0.5, 0.75], "rhs": {"type": "mul", "lhs": 2.0, "rhs": [0.1, 0.2, 0.3]}})))
    assert np.allclose(o.eval_texture2d(add, centres[:1]), [[0.45, 0.9, 1.35]], atol=1e-6)


def test_media_closed_forms(tmp_path, built):
    """Homogeneous Tr = exp(-sigma_a d) per channel (homogeneous.cpp:60-73) and constant-density ratio tracking (mean = exp(-sigma d))."""
    from btrc_b200 import scenes
    from oracle import front

    for het in (False, True):
        d = tmp_path / ("het" if het else "hom")
        path = scenes.make_medium_scene(str(d), width=16, height=16, spp=1, grid_res=8, heterogeneous=het)
        if het:  # constant density 2.0 so that the expectation is known
            with open(d / "density.txt", "w") as f:
                f.write("float 1 4 4 4\n" + " ".join(["2.0"] * 64) + "\n")
        o = front.OracleScene(path)
        assert o.has_medium()
        n = 20000
        rng = np.random.default_rng(3)
        o_med = np.zeros((n, 4), f32)
        o_med[:, :3] = np.array([0, 0.8, 0], f32) + rng.uniform(-0.1, 0.1, (n, 3))
        o_med[:, 3] = np.zeros(n, np.uint32).view(f32)  # medium id 0 = the cube's medium
        d_t1 = np.zeros((n, 4), f32)
        d_t1[:, :3] = [0.3, 0.0, 0.0]
        d_t1[:, 3] = 1.0
        tr = o.shadow_tr(o_med, d_t1)
        if not het:
            sigma_a = 2.5 * (1 - np.array([0.9, 0.85, 0.8]))
            assert np.allclose(tr, np.exp(-0.3 * sigma_a), rtol=1e-5)
        else:
            # the medium's local cube is scaled by 0.55: world length 0.3 crosses density 2.0 for 0.3 world units
            assert abs(tr[:, 0].mean() - np.exp(-2.0 * 0.3)) < 0.01
            assert np.array_equal(tr, o.shadow_tr(o_med, d_t1))  # keyed on the ray: deterministic
