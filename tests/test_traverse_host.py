"""The product's BVH builder and traversal function, compiled for the host by tests/host/host_traverse.cpp, against the
oracle: bit-identical closest hits (instance, primitive, t, u, v) and occlusion on primary, bounce and shadow ray batches.
Runs without a GPU; the GPU parity tests check the same function inside the sm_100a kernels."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from . import util
from .util import ROOT, f32


@pytest.fixture(scope="module", params=["traverse_bvh2", "traverse_bvh2_ifif"])
def ht(request):
    src = os.path.join(ROOT, "tests", "host", "host_traverse.cpp")
    out = os.path.join(ROOT, "tests", "host", "_libhosttrav_%s.so" % request.param)
    deps = [src, os.path.join(ROOT, "btrc_b200/csrc/host/bvh.cpp"), os.path.join(ROOT, "btrc_b200/csrc/host/bvh.h"),
            os.path.join(ROOT, "btrc_b200/csrc/device/traverse.cuh"), os.path.join(ROOT, "btrc_b200/csrc/device/lbvh.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-DHT_TRAVERSE=" + request.param, src, deps[1], "-o", out],
                       check=True)
    L = ctypes.CDLL(out)
    L.ht_build.restype = ctypes.c_void_p
    L.ht_build.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ht_build_lbvh.restype = ctypes.c_void_p
    L.ht_build_lbvh.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ht_free.argtypes = [ctypes.c_void_p]
    L.ht_num_nodes.argtypes = [ctypes.c_void_p]
    L.ht_closest.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
    L.ht_any.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
    return L


BUILDER = ["sah"]  # which of the product's builders _build uses: the host SAH builder or the device LBVH functions run on the host


@pytest.fixture(params=["sah", "lbvh"])
def builder(request):
    BUILDER[0] = request.param
    yield request.param
    BUILDER[0] = "sah"


def _build(L, orc):
    verts, ids = orc.triangles()
    verts = np.ascontiguousarray(verts, f32)
    ids = np.ascontiguousarray(ids, np.uint32)
    h = (L.ht_build_lbvh if BUILDER[0] == "lbvh" else L.ht_build)(verts.shape[0], verts.ctypes.data, ids.ctypes.data)
    assert h
    return h


def _closest(L, h, o_med, d_t1):
    o_med = np.ascontiguousarray(o_med, f32); d_t1 = np.ascontiguousarray(d_t1, f32)
    n = o_med.shape[0]
    ids = np.zeros((n, 2), np.uint32)
    tuv = np.zeros((n, 3), f32)
    L.ht_closest(h, n, o_med.ctypes.data, d_t1.ctypes.data, ids.ctypes.data, tuv.ctypes.data, None)
    return ids, tuv


def _check_closest(L, h, orc, o_med, d_t1):
    flag, det = orc.closest_hit(o_med, d_t1)
    ids, tuv = _closest(L, h, o_med, d_t1)
    hit = ((flag >> 30) & 1).astype(bool)
    assert np.array_equal(hit, ids[:, 0] != 0xFFFFFFFF)
    assert np.array_equal(flag[hit] & 0x3FFFFFFF, ids[hit, 0])
    assert np.array_equal(det[hit, 1], ids[hit, 1])
    assert np.array_equal(det[hit][:, [0, 2, 3]], tuv[hit].view(np.uint32))
    return int(hit.sum())


def _walk(orc, st, bounces, visit):
    """Follows the oracle's own path states for a few bounces, calling visit(o_med, d_t1, shadow) on every ray batch."""
    for _ in range(bounces):
        if st["o_med"].shape[0] == 0:
            break
        visit(st["o_med"], st["d_t1"], False)
        flag, detail = orc.closest_hit(st["o_med"], st["d_t1"])
        parts = []
        if orc.has_medium():
            st, flag, mo = orc.medium(st, flag, detail)
            parts.append({k: v[(mo["status"] & 1) != 0] for k, v in mo["next"].items()})
        sh = orc.shade(st, flag, detail)
        m = (sh["status"] & 2) != 0
        if m.any():
            visit(sh["sh_o_med"][m], sh["sh_d_t1"][m], True)
        parts.append({k: v[(sh["status"] & 1) != 0] for k, v in sh["next"].items()})
        st = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in parts[0]}


def test_golden_rays_cornell(ht, builder, built, tmp_path):
    """777,600 primary rays of lib/bvh's benchmark on the fan-triangulated Cornell box (tests/golden): same hits as the oracle."""
    from oracle import front

    path, _ = util.write_fan_cornell_scene(str(tmp_path))[:2]
    orc = front.OracleScene(path)
    h = _build(ht, orc)
    o_med, d_t1 = util.bvh_benchmark_rays()
    assert _check_closest(ht, h, orc, o_med, d_t1) > 700000
    ht.ht_free(h)


@pytest.mark.parametrize("kind", ["cornell", "materials"])
def test_bounce_and_shadow_rays(ht, builder, built, tmp_path, kind):
    from btrc_b200 import scenes
    from oracle import front

    if kind == "cornell":
        path = scenes.make_cornell(str(tmp_path), width=96, height=96, spp=2, state_count=20000)
    else:
        path = scenes.make_material_test(str(tmp_path), width=160, height=90, spp=1)
    orc = front.OracleScene(path)
    h = _build(ht, orc)
    assert ht.ht_num_nodes(h) > 0
    n_rays = [0, 0]

    def visit(o_med, d_t1, shadow):
        if shadow:
            occ = np.zeros(o_med.shape[0], np.uint8)
            o = np.ascontiguousarray(o_med, f32); d = np.ascontiguousarray(d_t1, f32)
            ht.ht_any(h, o.shape[0], o.ctypes.data, d.ctypes.data, occ.ctypes.data)
            assert np.array_equal(occ, orc.any_hit(o, d))
            n_rays[1] += o.shape[0]
        else:
            _check_closest(ht, h, orc, o_med, d_t1)
            n_rays[0] += o_med.shape[0]

    st = orc.generate(0, orc.width * orc.height)
    _walk(orc, st, 5, visit)
    assert n_rays[0] > orc.width * orc.height and n_rays[1] > 0
    ht.ht_free(h)


def test_edge_cases(ht, builder, built, tmp_path):
    """Axis-parallel rays (zero direction components), zero-length and negative intervals, rays starting on geometry."""
    from btrc_b200 import scenes
    from oracle import front

    path = scenes.make_cornell(str(tmp_path), width=32, height=32, spp=1, state_count=4096)
    orc = front.OracleScene(path)
    h = _build(ht, orc)
    rng = np.random.default_rng(5)
    n = 20000
    o = np.zeros((n, 4), f32)
    d = np.zeros((n, 4), f32)
    o[:, :3] = rng.uniform(-1.2, 2.2, (n, 3)).astype(f32)
    axis = rng.integers(0, 3, n)
    d[np.arange(n), axis] = rng.choice(np.array([-1, 1], f32), n)
    d[: n // 2, 3] = np.finfo(f32).max
    d[n // 2:, 3] = rng.uniform(0, 2, n - n // 2).astype(f32)
    d[::97, 3] = 0.0
    d[1::97, [0, 1, 2]] = f32(-0.0) * d[1::97, [0, 1, 2]] + d[1::97, [0, 1, 2]]
    _check_closest(ht, h, orc, o, d)
    ht.ht_free(h)


# ---- two-level layout (instanced scenes): host/bvh.h build_two_level_bvh + traverse.cuh traverse_two_level ---------------------
def _build_two_level(L, orc, path):
    """Hands the harness what Scene::commit hands the builder: object-space meshes, instance -> mesh, world -> object matrices
    (from the product's own loader through the C ABI, host side only) and the world-space triangles."""
    from btrc_b200 import api

    s = api.Scene(path)
    n_inst = s.info.n_instances
    inst_mesh = np.array([s.instance_geometry(i) for i in range(n_inst)], np.int32)
    n_mesh = int(inst_mesh.max()) + 1
    meshes = [np.ascontiguousarray(s.geometry_positions(g).reshape(-1, 9), f32) for g in range(n_mesh)]
    inv = np.ascontiguousarray(np.stack([s.instance_transform(i)[1].reshape(12) for i in range(n_inst)]), f32)
    mesh_ntri = np.array([m.shape[0] for m in meshes], np.int32)
    mesh_verts = np.ascontiguousarray(np.concatenate(meshes), f32)
    verts, ids = orc.triangles()
    verts = np.ascontiguousarray(verts, f32); ids = np.ascontiguousarray(ids, np.uint32)
    L.ht2_build.restype = ctypes.c_void_p
    L.ht2_build.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ht2_free.argtypes = [ctypes.c_void_p]
    L.ht2_info.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.ht2_closest.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
    L.ht2_any.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
    h = L.ht2_build(n_mesh, mesh_ntri.ctypes.data, mesh_verts.ctypes.data, n_inst, inst_mesh.ctypes.data, inv.ctypes.data, verts.shape[0],
                    verts.ctypes.data, ids.ctypes.data)
    s.close()
    assert h, "instance-major world triangle list does not match the meshes"
    return h


@pytest.mark.parametrize("kind", ["materials", "instanced", "cornell"])
def test_two_level_layout_matches_oracle(ht, built, tmp_path, kind):
    """Instance tree + one object-space tree per mesh, triangles tested in world space: closest hits (ids and t / u / v bits) and
    occlusion identical to the oracle on primary, bounce and shadow rays -- rotated / translated / scaled instances of shared
    meshes (15 spheres of one mesh; 40 instances of a blob under random rigid transforms and scales), and a scene without reuse."""
    from btrc_b200 import scenes
    from oracle import front

    if kind == "materials":
        path = scenes.make_material_test(str(tmp_path), width=160, height=90, spp=1, rings=24, segments=32)
    elif kind == "instanced":
        path = scenes.make_instanced_scene(str(tmp_path), width=160, height=90, spp=1, rings=21, segments=40, n_instances=40, n_lights=6, state_count=40000)
    else:
        path = scenes.make_cornell(str(tmp_path), width=96, height=96, spp=1, state_count=20000)
    orc = front.OracleScene(path)
    h = _build_two_level(ht, orc, path)
    assert ht.ht2_info(h, 1) > 0 and ht.ht2_info(h, 2) < 60
    n_rays = [0, 0]

    def visit(o_med, d_t1, shadow):
        o = np.ascontiguousarray(o_med, f32); d = np.ascontiguousarray(d_t1, f32)
        n = o.shape[0]
        if shadow:
            occ = np.zeros(n, np.uint8)
            ht.ht2_any(h, n, o.ctypes.data, d.ctypes.data, occ.ctypes.data)
            assert np.array_equal(occ, orc.any_hit(o, d))
            n_rays[1] += n
        else:
            flag, det = orc.closest_hit(o, d)
            ids = np.zeros((n, 2), np.uint32); tuv = np.zeros((n, 3), f32)
            ht.ht2_closest(h, n, o.ctypes.data, d.ctypes.data, ids.ctypes.data, tuv.ctypes.data, None)
            hit = ((flag >> 30) & 1).astype(bool)
            assert np.array_equal(hit, ids[:, 0] != 0xFFFFFFFF)
            assert np.array_equal(flag[hit] & 0x3FFFFFFF, ids[hit, 0]) and np.array_equal(det[hit, 1], ids[hit, 1])
            assert np.array_equal(det[hit][:, [0, 2, 3]], tuv[hit].view(np.uint32))
            n_rays[0] += n

    st = orc.generate(0, orc.width * orc.height)
    _walk(orc, st, 6, visit)
    assert n_rays[0] > orc.width * orc.height and n_rays[1] > 0
    ht.ht2_free(h)


# ---- compressed 8-wide layout ("bvh_layout": "wide8"): host/bvh.h build_wide_bvh8 + traverse.cuh traverse_bvh8 ---------------------
def _wide8(L, orc):
    L.ht8_build.restype = ctypes.c_void_p
    L.ht8_build.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.ht8_free.argtypes = [ctypes.c_void_p]
    L.ht8_info.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.ht8_closest.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
    L.ht8_any.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
    verts, ids = orc.triangles()
    verts = np.ascontiguousarray(verts, f32); ids = np.ascontiguousarray(ids, np.uint32)
    h = L.ht8_build(verts.shape[0], verts.ctypes.data, ids.ctypes.data)
    assert h
    return h, verts.shape[0]


def _check_closest8(L, h, orc, o_med, d_t1):
    o = np.ascontiguousarray(o_med, f32); d = np.ascontiguousarray(d_t1, f32)
    n = o.shape[0]
    flag, det = orc.closest_hit(o, d)
    ids = np.zeros((n, 2), np.uint32); tuv = np.zeros((n, 3), f32)
    steps = np.zeros(2, np.uint64)
    L.ht8_closest(h, n, o.ctypes.data, d.ctypes.data, ids.ctypes.data, tuv.ctypes.data, steps.ctypes.data)
    hit = ((flag >> 30) & 1).astype(bool)
    assert np.array_equal(hit, ids[:, 0] != 0xFFFFFFFF)
    assert np.array_equal(flag[hit] & 0x3FFFFFFF, ids[hit, 0]) and np.array_equal(det[hit, 1], ids[hit, 1])
    assert np.array_equal(det[hit][:, [0, 2, 3]], tuv[hit].view(np.uint32))
    return int(hit.sum()), steps


@pytest.mark.parametrize("kind", ["golden", "materials", "instanced", "edge"])
def test_wide8_layout_matches_oracle(ht, built, tmp_path, kind):
    """The 8-wide compressed tree (80-byte nodes, 8-bit child boxes in a per-node frame): closest hits (ids and t / u / v bits) and
    occlusion identical to the oracle -- lib/bvh's 777,600 benchmark rays on the fan-triangulated Cornell box; primary, bounce and
    shadow rays of the material-test and instanced scenes; axis-parallel rays, empty and zero-length intervals.  The quantised boxes
    must never cull a triangle the exact test accepts (they are widened by one quantum per side, host/bvh.cpp)."""
    if "ifif" in ht._name:
        pytest.skip("the wide walk does not depend on the binary loop form")
    from btrc_b200 import scenes
    from oracle import front

    if kind == "golden":
        path, _ = util.write_fan_cornell_scene(str(tmp_path))[:2]
        orc = front.OracleScene(path)
        h, n_tris = _wide8(ht, orc)
        o_med, d_t1 = util.bvh_benchmark_rays()
        n_hit, steps = _check_closest8(ht, h, orc, o_med, d_t1)
        assert n_hit > 700000
        ht.ht8_free(h)
        return
    if kind == "edge":
        path = scenes.make_cornell(str(tmp_path), width=32, height=32, spp=1, state_count=4096)
        orc = front.OracleScene(path)
        h, n_tris = _wide8(ht, orc)
        rng = np.random.default_rng(5)
        n = 20000
        o = np.zeros((n, 4), f32); d = np.zeros((n, 4), f32)
        o[:, :3] = rng.uniform(-1.2, 2.2, (n, 3)).astype(f32)
        axis = rng.integers(0, 3, n)
        d[np.arange(n), axis] = rng.choice(np.array([-1, 1], f32), n)
        d[: n // 2, 3] = np.finfo(f32).max
        d[n // 2:, 3] = rng.uniform(0, 2, n - n // 2).astype(f32)
        d[::97, 3] = 0.0
        _check_closest8(ht, h, orc, o, d)
        ht.ht8_free(h)
        return
    if kind == "materials":
        path = scenes.make_material_test(str(tmp_path), width=160, height=90, spp=1, rings=24, segments=32)
    else:
        path = scenes.make_instanced_scene(str(tmp_path), width=160, height=90, spp=1, rings=21, segments=40, n_instances=40, n_lights=6, state_count=40000)
    orc = front.OracleScene(path)
    h, n_tris = _wide8(ht, orc)
    # compact: 80 bytes per wide node against 64 per binary node, and several times fewer nodes
    assert 0 < ht.ht8_info(h, 0) * 80 < 0.5 * n_tris * 64 and ht.ht8_info(h, 1) <= 30
    n_rays = [0, 0]

    def visit(o_med, d_t1, shadow):
        o = np.ascontiguousarray(o_med, f32); d = np.ascontiguousarray(d_t1, f32)
        if shadow:
            occ = np.zeros(o.shape[0], np.uint8)
            ht.ht8_any(h, o.shape[0], o.ctypes.data, d.ctypes.data, occ.ctypes.data)
            assert np.array_equal(occ, orc.any_hit(o, d))
            n_rays[1] += o.shape[0]
        else:
            _check_closest8(ht, h, orc, o, d)
            n_rays[0] += o.shape[0]

    st = orc.generate(0, orc.width * orc.height)
    _walk(orc, st, 6, visit)
    assert n_rays[0] > orc.width * orc.height and n_rays[1] > 0
    ht.ht8_free(h)
