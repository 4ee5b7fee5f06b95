"""Oracle restatement of the volume primitives (core/volume.cpp, core/volume/{bvh,aggregate}.cpp) against closed forms.
The reference ships no fixtures for this path (parity unpinned by the reference: SURVEY 8c); the pins are the
Beer-Lambert law for constant densities (ratio tracking is unbiased), additivity of densities inside an overlap, and the
geometry of boundary crossings."""
import json
import os

import numpy as np
import pytest

from tests.util import f32


def _scene(tmp_path, volumes):
    from btrc_b200 import scenes

    path = scenes.make_volume_scene(str(tmp_path), width=16, height=16, spp=1, grid_res=8, state_count=1024)
    root = json.load(open(path))
    root["scene"]["volumes"] = volumes
    json.dump(root, open(path, "w"))
    return path


def _mean_tr(o, a, b, n=40000, medium_id=None):
    """Mean ratio-tracking transmittance of n shadow rays a -> b (t1 = 1) whose streams differ through a tiny jitter of `a`."""
    rng = np.random.default_rng(3)
    o_med = np.zeros((n, 4), f32)
    d_t1 = np.zeros((n, 4), f32)
    jit = rng.uniform(-1e-4, 1e-4, (n, 3)).astype(f32)
    o_med[:, :3] = np.asarray(a, f32) + jit
    mid = o.vol_prim_medium_id() if medium_id is None else medium_id
    o_med[:, 3] = np.array([mid], np.uint32).view(f32)[0]
    d_t1[:, :3] = np.asarray(b, f32) - o_med[:, :3]
    d_t1[:, 3] = 1.0
    return o.shadow_tr(o_med, d_t1)[:, 0].astype(np.float64).mean()


def test_beer_lambert_single_box(built, tmp_path):
    from oracle import front

    vols = [{"o": [-0.5, 0.5, -0.5], "x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1], "sigma_t": 2.0, "albedo": [0.8, 0.8, 0.8]}]
    o = front.OracleScene(_scene(tmp_path, vols))
    assert o.has_medium()
    # straight through the box along x: 1.0 inside
    tr = _mean_tr(o, (-0.9, 1.0, 0.0), (0.9, 1.0, 0.0))
    assert abs(tr - np.exp(-2.0 * 1.0)) < 4e-3
    # a chord that misses the box entirely
    assert abs(_mean_tr(o, (-0.9, 0.2, 0.0), (0.9, 0.2, 0.0)) - 1.0) < 1e-6
    # segment that ends inside the box: 0.4 inside
    assert abs(_mean_tr(o, (-0.9, 1.0, 0.0), (-0.1, 1.0, 0.0)) - np.exp(-2.0 * 0.4)) < 4e-3


def test_overlap_adds_densities(built, tmp_path):
    from oracle import front

    vols = [{"o": [-0.5, 0.5, -0.5], "x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1], "sigma_t": 2.0, "albedo": [0.8, 0.8, 0.8]},
            {"o": [0.0, 0.5, -0.5], "x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1], "sigma_t": 1.0, "albedo": [0.2, 0.9, 0.2]}]
    o = front.OracleScene(_scene(tmp_path, vols))
    # along x at y = 1: [-0.5, 0] density 2, [0, 0.5] density 3, [0.5, 1.0] density 1 (the ray stops at 0.9: 0.4 of it)
    want = np.exp(-(2.0 * 0.5 + 3.0 * 0.5 + 1.0 * 0.4))
    assert abs(_mean_tr(o, (-0.9, 1.0, 0.0), (0.9, 1.0, 0.0)) - want) < 4e-3


def test_sheared_box_and_scatter_rate(built, tmp_path):
    """A sheared parallelepiped: the inside test uses x/|x|^2 etc. as in volume/bvh.cpp:270-279, i.e. coordinates along the
    (non-orthogonal) edge directions.  Delta tracking scatters with probability 1 - exp(-sigma L)."""
    from oracle import front

    vols = [{"o": [-0.5, 0.5, -0.5], "x": [1, 0, 0], "y": [0, 1, 0], "z": [0, 0, 1], "sigma_t": 1.5, "albedo": [0.7, 0.7, 0.7]}]
    path = _scene(tmp_path, vols)
    o = front.OracleScene(path)
    n = 256
    st = o.generate(0, n)
    # camera rays through the image centre region traverse the box front to back (1.0 thick); count the scattered share
    flag, detail = o.closest_hit(st["o_med"], st["d_t1"])
    scat = 0
    tot = 0
    for rep in range(40):
        stt = o.generate(rep * n, n)
        fl, de = o.closest_hit(stt["o_med"], stt["d_t1"])
        st2, fl2, mo = o.medium(stt, fl, de)
        px = stt["pixel"]
        centre = (abs(px[:, 0].astype(int) - 8) <= 1) & (abs(px[:, 1].astype(int) - 8) <= 1)
        scat += int(((fl2 >> 31) & 1)[centre].sum())
        tot += int(centre.sum())
    assert tot > 300
    p = scat / tot
    assert abs(p - (1 - np.exp(-1.5 * 1.0))) < 0.08


def test_volume_known_answer_vectors(built, tmp_path):
    """Frozen transmittances (tests/golden/volumes_kat.json, written by tests/golden/make_golden_volumes.py)."""
    import importlib.util

    from oracle import front
    from tests.util import ROOT

    spec = importlib.util.spec_from_file_location("make_golden_volumes", os.path.join(ROOT, "tests", "golden", "make_golden_volumes.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "volumes_kat.json")))
    o = front.OracleScene(mg.scene(str(tmp_path)))
    o_med = np.array(kat["o_med_bits"], np.uint32).view(f32)
    d_t1 = np.array(kat["d_t1_bits"], np.uint32).view(f32)
    tr = o.shadow_tr(o_med, d_t1)
    assert tr.view(np.uint32).tolist() == kat["tr_bits"]
    assert tr.min() >= 0.0 and tr.max() <= 1.0 and len(np.unique(tr)) > 4  # constant boxes give 0 / 1 (density == majorant), the grid box fractions
