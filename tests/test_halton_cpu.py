"""Halton sampler (builtin/sampler/halton.cpp, utils/low_discrepancy.cpp): the oracle's restatement against the reference's own
prime table and against the defining properties of an Owen-scrambled Halton sequence.  The reference's wavefront renderer is
compiled with IndependentSampler (wavefront/common.h:36-37) and its low-discrepancy code is CUJ device code, so there is no
reference output to compare with: parity unpinned by the reference beyond the prime table."""
import ctypes
import json
import os

import numpy as np
import pytest

from tests.util import ROOT, f32

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbtrcref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libbtrcref.so not built (reference tree absent)")
def test_prime_table_matches_reference(built):
    from oracle import front

    R = ctypes.CDLL(REF_SO)
    L = front.lib()
    assert R.ref_prime_table_size() == 1000
    assert [L.orc_prime(i) for i in range(1000)] == [R.ref_prime(i) for i in range(1000)]


def test_scrambled_radical_inverse_is_stratified(built):
    """Owen scrambling permutes digits: the first b^m points of the base-b sequence put exactly one point in each interval of
    width b^-m, for every hash."""
    from oracle import front

    L = front.lib()
    for base_index, base, m in ((0, 2, 8), (1, 3, 5), (2, 5, 3), (10, 31, 2)):
        assert L.orc_prime(base_index) == base
        n = base ** m
        for h in (0, 1, 0x9E3779B97F4A7C15, 123456789):
            v = np.array([L.orc_owen_scrambled_radical_inverse(base_index, a, h) for a in range(n)], np.float64)
            assert v.min() >= 0 and v.max() < 1
            cells = np.floor(v * n + 1e-9).astype(np.int64)
            # fp32: the digits below the m-th are random and may round a value up to the next cell boundary (at most a few of n)
            assert len(set(cells.tolist())) >= n - 2 and np.bincount(cells, minlength=n).max() <= 2, (base, h)
    # different hashes give different point sets
    a = [L.orc_owen_scrambled_radical_inverse(3, i, 1) for i in range(64)]
    b = [L.orc_owen_scrambled_radical_inverse(3, i, 2) for i in range(64)]
    assert a != b


def _halton_scene(tmp_path, w=96, h=64, spp=8, sampler="halton"):
    from btrc_b200 import scenes

    path = scenes.make_cornell(str(tmp_path), width=w, height=h, spp=spp, state_count=8192)
    root = json.load(open(path))
    root["renderer"]["sampler"] = sampler
    out = os.path.join(str(tmp_path), "scene_%s.json" % sampler)
    json.dump(root, open(out, "w"))
    return out


def test_halton_index_lands_in_its_pixel(built, tmp_path):
    """HaltonSampler's constructor (halton.cpp:46-67): the (unscrambled) first two dimensions of sample `index` fall inside the
    pixel modulo 128, and consecutive samples of a pixel are one stride apart."""
    from oracle import front

    W, H, spp = 96, 64, 3
    o = front.OracleScene(_halton_scene(tmp_path, W, H, spp))
    st = o.generate(0, W * H * spp)
    idx = st["rng"][:, 0].astype(np.uint64)
    assert np.all(st["rng"][:, 1] == (2 + 3) << 1)  # dimension after the filter's two draws and the time sample: 2 + 3... box filter draws 2, time 1
    px, py = st["pixel"][:, 0].astype(np.int64), st["pixel"][:, 1].astype(np.int64)

    def radical_inverse(base, a):
        inv, r = 1.0 / base, np.zeros(a.shape, np.float64)
        f = inv
        a = a.copy()
        while a.max() > 0:
            r += (a % base) * f
            a //= base
            f *= inv
        return r

    sx, sy = 128, 81  # base_scales: smallest 2^k >= min(96, 128), 3^k >= min(64, 128)
    assert np.array_equal(np.floor(radical_inverse(2, idx.astype(np.int64)) * sx + 1e-9).astype(np.int64), px % 128 % sx)
    assert np.array_equal(np.floor(radical_inverse(3, idx.astype(np.int64)) * sy + 1e-9).astype(np.int64), py % 128 % sy)
    npix = W * H
    assert np.array_equal(idx[npix:2 * npix] - idx[:npix], np.full(npix, sx * sy, np.uint64))


def test_halton_render_agrees_with_independent(built, tmp_path):
    from oracle import front

    a = front.OracleScene(_halton_scene(tmp_path, 48, 48, 16, "halton")).render()
    b = front.OracleScene(_halton_scene(tmp_path, 48, 48, 16, "independent")).render()
    ma, mb = a["color"][..., :3].mean(), b["color"][..., :3].mean()
    assert np.isfinite(a["color"]).all()
    assert abs(ma - mb) / mb < 0.05
    assert a["stats"]["n_gen"] == b["stats"]["n_gen"]


def test_unknown_sampler_is_an_error(built, tmp_path):
    from btrc_b200 import api

    with pytest.raises(api.BtrcError, match="unknown sampler type: sobol"):
        api.Scene(_halton_scene(tmp_path, 8, 8, 1, "sobol"))


def test_halton_known_answer_vectors(built):
    """Frozen vectors (tests/golden/halton_kat.json, written by tests/golden/make_golden_halton.py)."""
    from oracle import front

    L = front.lib()
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "halton_kat.json")))
    assert [L.orc_prime(i) for i in (0, 1, 2, 10, 100, 500, 999)] == kat["primes_tail"]
    for e in kat["radical_inverse"]:
        v = f32(L.orc_owen_scrambled_radical_inverse(e["base_index"], int(e["a"]), int(e["hash"])))
        assert int(np.array([v], f32).view(np.uint32)[0]) == e["bits"], e
