"""Converged-image bar of BASELINE.json ("converged-image RMSE against a high-spp reference").

Fixture: the oracle's 8192-spp render of the config-1 Cornell box at 64x64 (tests/golden/cornell64_converged.npy, written by
tests/golden/make_golden_converged.py) and the RMSE the oracle's own 64 / 256 / 1024 spp renders have against it
(cornell64_converged.json).  Sample seeds depend on (pixel, sample index) only (sampler/independent.cpp:11-15), so a render at
S spp takes the first S samples of the fixture's 8192: its RMSE against the fixture is a deterministic number up to the fp32
noise of the implementation, and it falls like 1/sqrt(S).

CPU leg: the oracle reproduces its recorded RMSE (the fixture is what the generating script says it is).
GPU leg: the CUDA renderer, through the C ABI, reaches the same RMSE at the same sample counts - tolerance 10 % of the
recorded value (the two implementations take the same paths up to rare fp32 decision flips: relMSE < 1e-3 at equal spp in
tests/test_gpu_parity.py) - and its mean differs from the converged mean by less than 1 % at 1024 spp (oracle: 0.1 %).
"""
import json
import os

import numpy as np
import pytest

from tests.util import GOLDEN


def _fixture():
    meta = json.load(open(os.path.join(GOLDEN, "cornell64_converged.json")))
    ref = np.load(os.path.join(GOLDEN, "cornell64_converged.npy")).astype(np.float64)
    assert ref.shape == (meta["height"], meta["width"], 3)
    return meta, ref


def _rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - b) ** 2)))


def _scene(tmp_path, meta, spp):
    from btrc_b200 import scenes

    return scenes.make_cornell(str(tmp_path / f"s{spp}"), width=meta["width"], height=meta["height"], spp=spp, state_count=100000)


def test_oracle_converges_to_fixture(tmp_path, built):
    from oracle import front

    meta, ref = _fixture()
    assert abs(ref.mean() - meta["mean"]) < 1e-7
    got = {}
    for spp in (64, 256):
        img = front.OracleScene(_scene(tmp_path, meta, spp)).render()["color"][..., :3]
        got[spp] = _rmse(img, ref)
        assert abs(got[spp] - meta["oracle_rmse_at_spp"][str(spp)]) <= 1e-3 * meta["oracle_rmse_at_spp"][str(spp)], (spp, got[spp])
    # Monte Carlo rate: 4x the samples, about half the error (0.36 here: the first S samples are part of the fixture's own)
    assert got[256] < 0.6 * got[64]
    r = meta["oracle_rmse_at_spp"]
    assert r["1024"] < 0.6 * r["256"] < 0.36 * r["64"]


@pytest.mark.gpu
def test_gpu_render_converges_to_fixture(tmp_path, built):
    from btrc_b200 import api

    meta, ref = _fixture()
    for spp in (64, 256, 1024):
        s = api.Scene(_scene(tmp_path, meta, spp))
        color, _, _, stats = s.render()
        s.close()
        assert stats["n_gen"] == meta["width"] * meta["height"] * spp
        want = meta["oracle_rmse_at_spp"][str(spp)]
        got = _rmse(color[..., :3], ref)
        assert abs(got - want) <= 0.10 * want, (spp, got, want)
    assert abs(float(color[..., :3].astype(np.float64).mean()) - meta["mean"]) < 0.01 * meta["mean"]
