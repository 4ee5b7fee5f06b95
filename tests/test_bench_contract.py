"""bench.py keeps the driver's contract: ONE JSON line on stdout with the agreed keys, for the reference (CPU) arm here and for
our arm on the GPU box (short runs of the small Cornell workload)."""
import json
import os
import subprocess
import sys

import pytest

from tests.util import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]           # nothing but the result on stdout
    return json.loads(lines[0])


def test_reference_arm_line(built):
    d = _run(["--impl", "reference", "--workload", "c1_cornell_512_16spp", "--steps", "2", "--warmup", "1"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "c1_cornell_512_16spp" and d["steps"] == 2 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    from oracle import refcpu

    # "reference" = the reference's own sources built for the host (oracle/_ref/libbtrcrefcpu.so); "port" only where that is absent
    assert cb["kind"] == ("reference" if refcpu.available() else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "spp" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["ms_per_step"] > 0


@pytest.mark.gpu
def test_our_arm_line(built):
    d = _run(["--workload", "c1_cornell_512_16spp", "--steps", "2", "--warmup", "3", "--state-count", "1048576"], 900)
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "mrays_per_s"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "hbm" and rf["unit"] == "GB/s"
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= 512 * 512 * 16
    assert d["gpu_launches"] > 0 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
