"""CUDA shade stage against the REFERENCE'S OWN device code (oracle/_ref/libbtrcrefcpu.so, see oracle/refcpu/refcpu.cpp), on
identical batches, bucketed per material.

tests/test_reference_device_pins_cpu.py shows the oracle equal to that code bit for bit; here the comparison is made directly,
kernel against reference, so that no restatement stands between the two.  The reference library compiles a scene's kernels
with the host compiler when the scene is opened; on a box without one the oracle (bit-identical, pinned) stands in and the test
says so.

Tolerances per bucket (W7 of the round-1 review: one pooled outlier budget can hide a wrong lobe weight of a rarely sampled
branch).  A continuation state is matched with its reference twin through its RNG state (every branch of the stage draws the
same count, so the key is exact); a state that continues on one side only is a FLIP (a discrete decision -- lobe choice, the
Disney firefly clamp, black-fringe rejection, roulette -- taken differently under fp32 fast-math); a matched state whose
throughput differs by more than RTOL is an OUTLIER.
"""
import json

import numpy as np
import pytest

from tests.util import f32

pytestmark = pytest.mark.gpu

W2, H2 = 96, 54
RTOL = 5e-3


def _bucket_names(path):
    ents = json.load(open(path))["scene"]["entities"]

    def name(e):
        m = e["material"]
        t = m["type"]
        if e.get("light"):
            return "light"
        if t == "disney":
            if "normal" in m:
                return "disney_normal_mapped"
            if m.get("transmission", 0):
                return "disney_transmission"
            if m.get("clearcoat", 0) or m.get("sheen", 0):
                return "disney_clearcoat_sheen"
            return "disney_metallic_textured"
        return t

    return [name(e) for e in ents]


@pytest.fixture(scope="module", params=["materials", "extras"])
def trio(request, tmp_path_factory, built):
    from btrc_b200 import api, scenes
    from oracle import front

    d = tmp_path_factory.mktemp("vs_ref_" + request.param)
    if request.param == "materials":
        path = scenes.make_material_test(str(d), width=W2, height=H2, spp=8, rings=12, segments=16, tex_res=64, state_count=60000)
    else:  # black / invisible surfaces, texture graphs (mul, add, transform), mirror / border / point addressing, box filter + tile indexing
        from tests.test_reference_device_pins_cpu import _make_extras

        path = _make_extras(str(d))
    s = api.Scene(path)
    o = front.OracleScene(path)
    ref, kind = o, "oracle (reference host build unavailable here)"
    try:
        from oracle import refcpu

        if refcpu.available():
            front.lib().orc_set_reduce_order(1)   # alias tables of the reference as built with libstdc++ (see the CPU pin test)
            o = front.OracleScene(path)
            ref, kind = refcpu.ReferenceScene(path), "reference"
    except Exception as e:  # no host compiler on the box
        kind = "oracle (reference host build failed: %s)" % str(e).splitlines()[-1][:80]
    yield s, o, ref, kind, path
    s.close()
    front.lib().orc_set_reduce_order(0)


def test_shade_against_reference_per_material(trio):
    s, o, ref, kind, path = trio
    names = _bucket_names(path)
    print("\nreference side:", kind)
    stats = {}
    st = o.generate(0, min(o.width * o.height * o.spp, 40000))
    for bounce in range(5):
        flag, detail = o.closest_hit(st["o_med"], st["d_t1"])
        want = ref.shade(st, flag, detail)
        if hasattr(ref, "to_oracle_ids"):
            ref.to_oracle_ids(want["next"]["o_med"])
        s.film_clear()
        nxt, shd = s.shade(st, flag, detail)
        hit = ((flag >> 30) & 1) != 0
        inst = flag & 0x3fffffff
        bucket = np.array([names[i] if h else "miss" for i, h in zip(inst, hit)], dtype=object)
        keep = (want["status"] & 1) != 0
        # input index of every continuation state, through the pixel + RNG key
        key_in = {}
        wn = want["next"]
        for j in np.nonzero(keep)[0]:
            key_in[(int(wn["rng"][j, 0]), int(wn["rng"][j, 1]))] = j
        got = {(int(a), int(b)): i for i, (a, b) in enumerate(nxt["rng"])}
        for name in set(bucket):
            stats.setdefault(name, dict(n=0, flips=0, matched=0, outliers=0, dir_out=0))
        for name in set(bucket):
            stats[name]["n"] += int((bucket == name).sum())
        matched_in, matched_out = [], []
        for k, j in key_in.items():
            if k in got:
                matched_in.append(j); matched_out.append(got[k])
            else:
                stats[bucket[j]]["flips"] += 1
        # continuations the kernel emitted that the reference did not: attribute them through the pixel of the state
        extra = [i for k, i in got.items() if k not in key_in]
        if extra:
            px_in = {(int(a), int(b)): idx for idx, (a, b) in enumerate(st["pixel"])}
            for i in extra:
                j = px_in.get((int(nxt["pixel"][i, 0]), int(nxt["pixel"][i, 1])))
                stats[bucket[j] if j is not None else "miss"]["flips"] += 1
        mi, mo = np.array(matched_in, np.int64), np.array(matched_out, np.int64)
        if mi.size:
            assert np.array_equal(nxt["pixel"][mo], wn["pixel"][mi])
            assert np.array_equal(nxt["beta_depth"][mo, 3].view(np.uint32), wn["beta_depth"][mi, 3].view(np.uint32))
            a, b = nxt["beta_depth"][mo, :3].astype(np.float64), wn["beta_depth"][mi, :3].astype(np.float64)
            bad = (np.abs(a - b) > RTOL * np.abs(b) + 1e-5).any(axis=1)
            dbad = (np.abs(nxt["d_t1"][mo, :3] - wn["d_t1"][mi, :3]) > 2e-3).any(axis=1)
            for name in set(bucket[mi]):
                m = bucket[mi] == name
                stats[name]["matched"] += int(m.sum()); stats[name]["outliers"] += int(bad[m].sum()); stats[name]["dir_out"] += int(dbad[m].sum())
        # next bounce from the reference's continuation states
        st = {k: np.ascontiguousarray(v[keep]) for k, v in wn.items()}
        if st["o_med"].shape[0] < 64:
            break
    print("bucket                        states  matched   flips  flip%   throughput-outliers%  direction-outliers%")
    for name, v in sorted(stats.items()):
        print("%-28s %7d  %7d  %6d  %5.2f   %8.2f   %8.2f" % (name, v["n"], v["matched"], v["flips"], 100.0 * v["flips"] / max(v["n"], 1),
                                                         100.0 * v["outliers"] / max(v["matched"], 1), 100.0 * v["dir_out"] / max(v["matched"], 1)))
    # budgets per bucket: flips and outliers as fractions of the bucket's own states
    # Measured on B200 against the reference's own code (profiles/r2d_pytest_gpu.log): 3 flips in 46,000 states (all on the
    # normal-mapped ground); throughput outliers 1.3 % clearcoat (gloss 0.9: a GGX lobe with alpha ~ 1e-2 whose D term moves by
    # percents under a 1e-7 change of the half vector), 0.7 % / 1.6 % for the textured buckets (8-bit texture filter weights), 0.07 %
    # transmission, none for glass / metal / mirror.  Budgets: about twice the measured rates, per bucket.
    flip_budget = {"default": 0.001}
    outlier_budget = {"default": 0.003, "disney_normal_mapped": 0.03, "disney_metallic_textured": 0.015, "disney_clearcoat_sheen": 0.03}
    if "vs_ref_extras" in path:  # this scene textures its diffuse and metal surfaces (texture graph over 32^2 images, point / mirror / border addressing)
        outlier_budget.update({"diffuse": 0.01, "metal": 0.01})   # measured 0.29 % / 0.45 %
    for name, v in stats.items():
        if v["n"] < 200 or name in ("miss", "light"):
            continue
        assert v["flips"] <= flip_budget.get(name, flip_budget["default"]) * v["n"] + 2, (name, v)
        assert v["outliers"] <= outlier_budget.get(name, outlier_budget["default"]) * max(v["matched"], 1) + 2, (name, v)
        assert v["dir_out"] <= 0.004 * max(v["matched"], 1) + 2, (name, v)
    assert sum(v["matched"] for v in stats.values()) > 5000


def test_byte_grid_medium_matches_oracle(tmp_path, built):
    """"byte" text grids (SNorm8 3-D arrays, utils/cuda/array.cpp:331-355 -- refused in round 1): the medium stage on identical
    batches and the whole render against the oracle, whose decoding of such grids is pinned to the reference's own loader bit for
    bit (tests/test_reference_device_pins_cpu.py, kind "medium_bytes")."""
    from btrc_b200 import api
    from oracle import front
    from tests.test_reference_device_pins_cpu import _make_medium_bytes

    path = _make_medium_bytes(str(tmp_path))
    s, o = api.Scene(path), front.OracleScene(path)
    assert s.info.has_medium
    st = o.generate(0, 64 * 64 * 2)
    flag, detail = o.closest_hit(st["o_med"], st["d_t1"])
    # second bounce: rays that start around the medium
    sh = o.shade(*o.medium(st, flag, detail)[:2], detail)
    keep = (sh["status"] & 1) != 0
    st = {k: np.ascontiguousarray(v[keep]) for k, v in sh["next"].items()}
    flag, detail = o.closest_hit(st["o_med"], st["d_t1"])
    so, fo, mo = o.medium(st, flag, detail)
    sg, fg, ng, shg = s.medium(st, flag, detail)
    scat_o, scat_g = (fo >> 31) != 0, (fg >> 31) != 0
    assert scat_o.mean() > 0.05
    # a collision decision compares a random number with density / majorant, both through the texture unit: rare flips only
    assert (scat_o != scat_g).mean() < 0.004
    same = scat_o == scat_g
    assert np.array_equal(sg["o_med"][same, 3].view(np.uint32), so["o_med"][same, 3].view(np.uint32))          # resolved medium ids
    quiet = same & ~scat_o
    assert np.array_equal(sg["rng"][quiet], so["rng"][quiet])                                                   # same number of draws
    color, _, _, stats = s.render()
    ref = o.render()
    for k in ("n_closest", "n_hit", "n_next", "n_shadow"):
        assert abs(stats[k] - ref["stats"][k]) <= 1e-2 * ref["stats"][k], (k, stats[k], ref["stats"][k])
    a, b = color[..., :3], ref["color"][..., :3]
    assert abs(a.mean() - b.mean()) < 0.02 * b.mean()
    s.close()
