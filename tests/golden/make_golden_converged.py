#!/usr/bin/env python
"""Freezes the converged-image fixture of BASELINE.json's second correctness bar ("converged-image RMSE against a high-spp
reference"): the oracle's render of the config-1 Cornell box at 64x64 / 8192 spp (tests/golden/cornell64_converged.npy,
float32 RGB) and, beside it, the RMSE of the oracle's own 64 / 256 / 1024 spp renders against that image
(cornell64_converged.json) - the noise floor a correct renderer reaches at those sample counts with the same sampler seeds.
About one minute of CPU on 8 cores."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from btrc_b200 import scenes  # noqa: E402
from oracle import front  # noqa: E402

W = H = 64
REF_SPP = 8192
CHECK_SPP = (64, 256, 1024)


def scene(d, spp):
    return scenes.make_cornell(d, width=W, height=H, spp=spp, state_count=100000)


def rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


if __name__ == "__main__":
    import __graft_entry__ as ge

    ge.build()
    with tempfile.TemporaryDirectory() as d:
        ref = front.OracleScene(scene(d, REF_SPP)).render()["color"][..., :3].astype(np.float32)
        at = {str(s): rmse(front.OracleScene(scene(d, s)).render()["color"][..., :3], ref) for s in CHECK_SPP}
    g = os.path.join(ROOT, "tests", "golden")
    np.save(os.path.join(g, "cornell64_converged.npy"), ref)
    json.dump({"width": W, "height": H, "ref_spp": REF_SPP, "mean": float(ref.astype(np.float64).mean()), "oracle_rmse_at_spp": at},
              open(os.path.join(g, "cornell64_converged.json"), "w"), indent=1)
    print("mean", ref.mean(), "rmse", at)
