"""One GPU parity test per BASELINE.json configuration AT ITS STATED GEOMETRY (image size, triangle counts, texture and grid
resolutions, depth limits); only the sample count is reduced so that the oracle side stays in seconds.

Every camera path is a function of (pixel, sample index) alone and splats into its own pixel only (generate.cpp:62-68
keeps filter-importance-sampled positions inside the pixel), so a pixel of the raw film is the sum over that pixel's paths: the
CUDA render of the whole image is compared with the oracle on a strided pixel subset (`orc_render_pixels`), pixel by pixel, and
the size-independent properties (weights, linearity over sample ranges and over tiles) are checked on all pixels.

Per-pixel bar.  Integer outcomes (weights) are exact.  Radiance sums agree to fp32 shading accuracy (the kernels use the
reference's fast-math contract, the oracle IEEE arithmetic) unless a discrete decision of one of the pixel's paths flips (lobe
choice, roulette against a threshold, an alias-table cell): such a pixel is an OUTLIER.  Outliers are counted per bucket of the
primary hit (material class) -- a wrong lobe weight in one material cannot hide in the average of the others -- and bounded per
bucket; the image-level error of the subset is bounded too.
"""
import json

import numpy as np
import pytest

from tests.util import f32

pytestmark = pytest.mark.gpu

REL_TOL = 2e-3       # per-pixel relative difference that still counts as "same paths" (fp32 fast-math vs IEEE arithmetic)
REL_TOL_TEXTURED = 1e-2  # scenes with image textures: the texture unit filters with 8-bit weights (1e-3 typical, 6e-3 at the 99th
                         # percentile per fetch against the oracle's exact weights), and most paths meet the textured ground
OUTLIER_BUDGET = {   # fraction of subset pixels per bucket whose radiance differs by more than the tolerance (2 spp, depth <= 10 / 16)
    "default": 0.02,
}


def _subset(npix, stride, phase=11):
    return np.arange(phase, npix, stride, dtype=np.uint32)


def _raw_film(s, begin=None, end=None):
    import torch

    i = s.info
    s.film_clear()
    if begin is None:
        begin, end = 0, i.width * i.height * i.spp
    st = s.render_range(begin, end)
    rad = torch.zeros(i.height, i.width, 4, device="cuda")
    wgt = torch.zeros(i.height, i.width, device="cuda")
    s.film_export(rad, wgt)
    return rad.cpu().numpy().reshape(-1, 4)[:, :3], wgt.cpu().numpy().reshape(-1), st


def _primary_buckets(o, path, idx):
    """Bucket name of every subset pixel: material type (+ a coarse flavour) of what the pixel's first camera ray hits."""
    root = json.load(open(path))
    ents = root["scene"]["entities"]

    def name(e):
        m = e["material"]
        t = m["type"]
        if e.get("light"):
            return "light"
        if t == "disney":
            if isinstance(m.get("color"), dict) and "normal" in m:
                return "disney_normal_mapped"
            if m.get("transmission", 0):
                return "disney_transmission"
            if m.get("clearcoat", 0) or m.get("sheen", 0):
                return "disney_clearcoat_sheen"
            return "disney_metallic_textured"
        return t

    names = [name(e) for e in ents]
    W, H = o.width, o.height
    out = np.empty(idx.shape[0], dtype=object)
    # sample 0 of pixel p is path index p (uniform mode) or p * spp (tile mode): generate row by row around the subset
    st = o.generate(0, W * H) if not root["renderer"].get("tile") else None
    assert st is not None
    flag, _ = o.closest_hit(st["o_med"][idx], st["d_t1"][idx])
    hit = (flag >> 30) & 1
    inst = flag & 0x3fffffff
    for k in range(idx.shape[0]):
        out[k] = names[inst[k]] if hit[k] else "miss"
    return out


def _compare_subset(rad, wgt, ref, buckets, spp, what, budget=None, image_tol=2e-2, rel_tol=REL_TOL, median_tol=1e-4):
    """rad, wgt: GPU raw film at the subset pixels; ref: oracle (n, 4) = radiance sum + weight."""
    budget = budget or {}
    assert np.array_equal(wgt, ref[:, 3]) and (wgt == spp).all(), what + ": film weights"
    a, b = rad.astype(np.float64), ref[:, :3].astype(np.float64)
    assert np.isfinite(a).all(), what
    scale = np.maximum(np.abs(b).max(axis=1), 1e-3 * max(b.mean(), 1e-6))
    rel = np.abs(a - b).max(axis=1) / scale
    report = {}
    for name in sorted(set(buckets)):
        m = buckets == name
        frac = float((rel[m] > rel_tol).mean())
        report[name] = (int(m.sum()), frac, float(np.median(rel[m])))
    print("\n%s: outliers per primary-hit bucket (pixels, fraction > %.0e, median rel. diff):" % (what, rel_tol))
    for name, (cnt, frac, med) in report.items():
        print("   %-28s %7d  %.4f  %.2e" % (name, cnt, frac, med))
    for name, (cnt, frac, med) in report.items():
        if cnt < 50:
            continue
        assert frac <= budget.get(name, budget.get("default", OUTLIER_BUDGET["default"])), (what, name, cnt, frac)
        assert med <= median_tol, (what, name, med)  # the typical pixel agrees to fp32 shading (/ texture filtering) accuracy
    # image-level: the subset's mean radiance and its relative MSE (outliers included: flipped paths are still valid samples)
    assert abs(a.mean() - b.mean()) <= 5e-3 * b.mean() + 1e-7, (what, a.mean(), b.mean())
    rel_mse = ((a - b) ** 2).mean() / (b.mean() ** 2 + 1e-12)
    assert rel_mse < image_tol, (what, rel_mse)
    return report


def test_c1_cornell_512_16spp(cornell_c1):
    """configs[0] in full: 512 x 512, 16 spp, every pixel against the oracle (4.2 M paths)."""
    from btrc_b200 import api
    from oracle import front

    s, o = api.Scene(cornell_c1), front.OracleScene(cornell_c1)
    rad, wgt, st = _raw_film(s)
    idx = np.arange(512 * 512, dtype=np.uint32)
    ref, rst = o.render_pixels(idx)
    assert st["n_gen"] == rst["n_gen"] == 512 * 512 * 16
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(st[k] - rst[k]) <= 1e-3 * rst[k], (k, st[k], rst[k])
    buckets = np.full(idx.shape[0], "diffuse", dtype=object)
    _compare_subset(rad, wgt, ref, buckets, 16, "C1", budget={"diffuse": 0.001}, image_tol=1e-3)   # measured: 0.0001
    s.close()


def test_c2_disney_1080p_full_geometry(tmp_path, built):
    """configs[1] at 1920 x 1080 with the 19,800-triangle spheres and 1024^2 textures; 2 of its 256 spp."""
    from btrc_b200 import api, scenes
    from oracle import front

    path = scenes.make_material_test(str(tmp_path), spp=2, state_count=4194304)
    s, o = api.Scene(path), front.OracleScene(path)
    assert s.info.width == 1920 and s.info.height == 1080 and s.info.n_triangles == o.num_triangles() > 297000
    rad, wgt, st = _raw_film(s)
    assert (wgt == 2).all()
    idx = _subset(1920 * 1080, 61)
    ref, _ = o.render_pixels(idx)
    buckets = _primary_buckets(o, path, idx)
    _compare_subset(rad[idx], wgt[idx], ref, buckets, 2, "C2", rel_tol=REL_TOL_TEXTURED, median_tol=3e-3,
                    budget={"default": 0.02})   # measured (profiles/r2d_pytest_gpu.log): at most 0.9 % in any bucket
    s.close()


def test_c3_medium_1080p_full_grid(tmp_path, built):
    """configs[2] at 1920 x 1080 with the 128^3 density grid (delta / ratio tracking, NEE through the medium); 2 spp."""
    from btrc_b200 import api, scenes
    from oracle import front

    path = scenes.make_medium_scene(str(tmp_path), spp=2, state_count=4194304)
    root = json.load(open(path))
    assert root["width"] == 1920 and root["height"] == 1080
    s, o = api.Scene(path), front.OracleScene(path)
    assert s.info.has_medium
    rad, wgt, st = _raw_film(s)
    assert (wgt == 2).all()
    idx = _subset(1920 * 1080, 61)
    ref, _ = o.render_pixels(idx)
    buckets = _primary_buckets(o, path, idx)
    # the density grid is a float 3-D texture: fetched exactly (the 8-bit filter weights apply to the interpolation, and both sides
    # see the same lattice), and the shadow-ray transmittance streams are keyed on the ray in oracle and kernels alike
    _compare_subset(rad[idx], wgt[idx], ref, buckets, 2, "C3", budget={"default": 0.005}, image_tol=1e-2)   # measured: no outlier at all
    s.close()


def test_c4_instanced_10m_full_geometry(tmp_path, built):
    """configs[3]: 250 x 40,000 triangles + 64 mesh lights, max depth 16, 1920 x 1080; 1 spp."""
    from btrc_b200 import api, scenes
    from oracle import front

    path = scenes.make_instanced_scene(str(tmp_path), spp=1, state_count=4194304)
    s, o = api.Scene(path), front.OracleScene(path)
    assert s.info.n_triangles == o.num_triangles() > 10_000_000 and s.info.n_lights == 64 and s.info.max_depth == 16
    rad, wgt, st = _raw_film(s)
    assert (wgt == 1).all()
    idx = _subset(1920 * 1080, 43)
    ref, _ = o.render_pixels(idx)
    buckets = _primary_buckets(o, path, idx)
    # measured: 1.7 % (diffuse) / 2.6 % (metal) of the pixels differ by more than 2e-3 -- paths are up to 16 bounces deep and most
    # of them meet a GGX lobe whose value is ill-conditioned in fp32; medians 7e-7 / 1.3e-6
    _compare_subset(rad[idx], wgt[idx], ref, buckets, 1, "C4", budget={"default": 0.05}, image_tol=5e-2)
    s.close()


def test_c5_4k_two_rank_partitions(tmp_path, built):
    """configs[4] at 3840 x 2160: the film of a 2-rank SAMPLE split and of a 2-rank TILE split (interleaved 256 x 16 tiles) each sums to
    the single render (what the NCCL reduce computes), and the single render matches the oracle on a pixel subset; 2 spp."""
    from btrc_b200 import api, partition, scenes
    from oracle import front

    W, H, spp = 3840, 2160, 2
    path = scenes.make_material_test(str(tmp_path / "samples"), width=W, height=H, spp=spp, state_count=8388608)
    s = api.Scene(path)
    npix = W * H
    full, wfull, _ = _raw_film(s)
    assert (wfull == spp).all()
    parts = [_raw_film(s, *partition.path_range(r, 2, spp, npix)) for r in range(2)]
    assert np.array_equal(parts[0][1] + parts[1][1], wfull)
    ssum = parts[0][0] + parts[1][0]
    # same paths, same splats; only the order of the float additions inside a pixel differs
    assert np.abs(ssum - full).max() <= 1e-5 * max(1.0, float(full.max()))
    o = front.OracleScene(path)
    idx = _subset(npix, 211)
    ref, _ = o.render_pixels(idx)
    _compare_subset(full[idx], wfull[idx], ref, _primary_buckets(o, path, idx), spp, "C5", rel_tol=REL_TOL_TEXTURED, median_tol=3e-3,
                    budget={"default": 0.025})  # measured: at most 1.1 % in any bucket
    s.close()

    # tile split: tile-mode indexing (generate.cpp:56-60), every rank renders all samples of its tiles
    tpath = scenes.make_material_test(str(tmp_path / "tiles"), width=W, height=H, spp=spp, state_count=8388608)
    root = json.load(open(tpath)); root["renderer"]["tile"] = True; json.dump(root, open(tpath, "w"))
    t = api.Scene(tpath)
    tfull, twfull, _ = _raw_film(t)
    import torch

    def part_film(r):
        t.film_clear()
        t.render_part(2, r)                          # what every rank of `btrc --ranks 2` / bench.py --gpus 2 calls
        rad = torch.zeros(H, W, 4, device="cuda"); wgt = torch.zeros(H, W, device="cuda")
        t.film_export(rad, wgt)
        return rad.cpu().numpy().reshape(-1, 4)[:, :3], wgt.cpu().numpy().reshape(-1)

    acc, wacc = np.zeros_like(tfull), np.zeros_like(twfull)
    owner = partition.tile_owner(2, W, H).reshape(-1)  # 256 x 16 tiles at this geometry (15 per row), tile k of the row-major tile order -> rank k % 2
    assert partition.tiles_are_2d(2, W, H, partition.strip_rows(2, H)) and partition.tile_width(2, W) == 256
    tol = 1e-5 * max(1.0, float(tfull.max()))
    for r in range(2):
        prad, pwgt = part_film(r)
        own = owner == r
        # the rank renders every sample of exactly its own tiles: the same (pixel, sample) paths as the single render there
        assert np.array_equal(pwgt[own], twfull[own]) and not pwgt[~own].any() and not prad[~own].any()
        assert np.abs(prad[own] - tfull[own]).max() <= tol
        acc += prad; wacc += pwgt
    assert np.array_equal(wacc, twfull)
    assert np.abs(acc - tfull).max() <= 1e-5 * max(1.0, float(tfull.max()))
    # tile-mode and uniform-mode indexing enumerate the same (pixel, sample) pairs: same image up to addition order
    assert np.abs(tfull - full).max() <= 1e-5 * max(1.0, float(full.max()))
    t.close()


@pytest.mark.parametrize("tile", [False, True])
def test_partitions_with_ragged_edges(tmp_path, built, tile):
    """The closed-form film weights and the partitions on sizes that divide nothing: an image whose height is no multiple of the strip
    height split over 3 parts (tile mode), sample ranges over 3 parts of 7 spp (uniform mode), and arbitrary index ranges that cut
    through sample planes / pixels -- parts sum to the whole render, weights exactly."""
    import torch

    from btrc_b200 import api, scenes

    W, H, spp = 150, 77, 7
    path = scenes.make_cornell(str(tmp_path), width=W, height=H, spp=spp, state_count=30000)
    if tile:
        root = json.load(open(path)); root["renderer"]["tile"] = True; json.dump(root, open(path, "w"))
    s = api.Scene(path)
    full, wfull, _ = _raw_film(s)
    assert (wfull == spp).all()

    def film_of(call):
        s.film_clear()
        call()
        rad = torch.zeros(H, W, 4, device="cuda"); wgt = torch.zeros(H, W, device="cuda")
        s.film_export(rad, wgt)
        return rad.cpu().numpy().reshape(-1, 4)[:, :3], wgt.cpu().numpy().reshape(-1)

    tol = 1e-5 * max(1.0, float(full.max()))
    for rows in (0, 5):
        acc, wacc = np.zeros_like(full), np.zeros_like(wfull)
        for part in range(3):
            rad, wgt = film_of(lambda: s.render_part(3, part, rows))
            if tile:
                assert set(np.unique(wgt)) <= {0.0, float(spp)}            # a part owns whole pixels
            acc += rad; wacc += wgt
        assert np.array_equal(wacc, wfull) and np.abs(acc - full).max() <= tol
    total = W * H * spp
    cuts = [0, 12345, total // 2 + 3, total - 1, total]
    acc, wacc = np.zeros_like(full), np.zeros_like(wfull)
    for b, e in zip(cuts[:-1], cuts[1:]):
        rad, wgt = film_of(lambda: s.render_range(b, e))
        acc += rad; wacc += wgt
    assert np.array_equal(wacc, wfull) and np.abs(acc - full).max() <= tol
    s.close()


@pytest.mark.parametrize("spp", [7, 64])
def test_partitions_with_2d_tiles(tmp_path, built, spp):
    """btrc_render_part on an image that divides into 2-D tiles (64 wide for 2 and 8 parts, 3 tiles per row; 64 wide for 3 parts: no width is coprime) (the C5 case in small): every part renders exactly the pixels
    partition.tile_owner gives it, all samples of them (7 spp: pixel-major inside a tile; 64 spp: warps of 32 samples of one pixel,
    pixel after pixel), and the parts sum to the single render."""
    import torch

    from btrc_b200 import api, partition, scenes

    W, H = 192, 96
    path = scenes.make_cornell(str(tmp_path), width=W, height=H, spp=spp, state_count=50000)
    root = json.load(open(path)); root["renderer"]["tile"] = True; json.dump(root, open(path, "w"))
    s = api.Scene(path)
    full, wfull, _ = _raw_film(s)
    assert (wfull == spp).all()
    tol = 1e-5 * max(1.0, float(full.max()))
    for parts in (2, 3, 8):
        rows = partition.strip_rows(parts, H)
        assert partition.tiles_are_2d(parts, W, H, rows)
        owner = partition.tile_owner(parts, W, H).reshape(-1)
        acc, wacc = np.zeros_like(full), np.zeros_like(wfull)
        for part in range(parts):
            s.film_clear()
            s.render_part(parts, part)
            rad = torch.zeros(H, W, 4, device="cuda"); wgt = torch.zeros(H, W, device="cuda")
            s.film_export(rad, wgt)
            rad = rad.cpu().numpy().reshape(-1, 4)[:, :3]; wgt = wgt.cpu().numpy().reshape(-1)
            own = owner == part
            assert np.array_equal(wgt[own], wfull[own]) and not wgt[~own].any() and not rad[~own].any()
            assert np.abs(rad[own] - full[own]).max() <= tol
            acc += rad; wacc += wgt
        assert np.array_equal(wacc, wfull) and np.abs(acc - full).max() <= tol
    s.close()


def test_path_indices_beyond_32_bits(tmp_path, built):
    """Renders with more than 2^32 camera paths (BASELINE config 5 on one GPU has 8.5e9): the generate kernel divides the path index
    in 32 bits when the whole range fits and in 64 bits otherwise.  The film weights are closed-form (weight_fill_kernel), so the
    witness of WHERE the kernel started its paths is the albedo plane: every camera path that hits adds its surface albedo to its
    own pixel (shade.cpp:118-126), i.e. albedo sum / albedo of the pixel = number of paths generated there.
    (a) index ranges that straddle 2^32 land in exactly the pixels the index space says (uniform: index = sample * W*H + pixel,
    generate.cpp:49-55; tile: index = pixel * spp + sample, :56-60); (b) uniform mode: a whole render of 4.32e9 paths in the
    block-coherent order generates every (pixel, sample) once."""
    import torch

    from btrc_b200 import api, scenes

    W = H = 2048
    spp = 1030
    npix = W * H
    assert npix * spp > 2 ** 32

    def planes(s, begin, end):
        s.film_clear()
        st = s.render_range(begin, end)
        rad = torch.zeros(H, W, 4, device="cuda"); wgt = torch.zeros(H, W, device="cuda"); alb = torch.zeros(H, W, 4, device="cuda")
        s.film_export(rad, wgt, alb)
        return wgt.cpu().numpy().reshape(-1), alb.cpu().numpy().reshape(-1, 4)[:, 0], st

    scn = {}
    for tile in (False, True):
        path = scenes.make_cornell(str(tmp_path / ("t%d" % tile)), width=W, height=H, spp=spp, state_count=33554432, min_depth=1, max_depth=1, albedo=True, tile=tile)
        scn[tile] = api.Scene(path)
    _, alb0, _ = planes(scn[False], 0, npix)          # sample 0 of every pixel: the pixel's albedo (red channel)
    flat = alb0 > 0.01                                # pixels that see a surface
    assert flat.mean() > 0.9
    begin, end = 2 ** 32 - 3 * npix - 70001, 2 ** 32 + 2 * npix + 50003
    idx = np.arange(begin, end, dtype=np.int64)
    for tile in (False, True):
        pix = idx // spp if tile else idx % npix
        expect = np.bincount(pix, minlength=npix).astype(np.float64)
        wgt, alb, st = planes(scn[tile], begin, end)
        assert st["n_gen"] == end - begin and np.array_equal(wgt, expect.astype(np.float32))
        assert not alb[expect == 0].any()                                           # nothing started outside the range's pixels
        on = flat & (expect > 0)
        ratio = alb[on] / expect[on]
        # a pixel's albedo is constant except across material edges (sub-pixel jitter): all but a few per cent agree to 1e-3
        assert (np.abs(ratio - alb0[on]) <= 1e-3 * np.maximum(alb0[on], 0.05)).mean() > 0.97, tile
        # and the total is the number of paths times the mean albedo: a path started in a wrong pixel of equal albedo would still show above
        assert abs(alb[on].sum() / (expect[on] * alb0[on]).sum() - 1) < 2e-3
    wgt, alb, st = planes(scn[False], 0, npix * spp)
    assert st["n_gen"] == npix * spp and (wgt == spp).all()
    ratio = alb[flat] / spp
    assert (np.abs(ratio - alb0[flat]) <= 2e-3 * np.maximum(alb0[flat], 0.05)).mean() > 0.97
    assert abs(alb[flat].sum() / (spp * alb0[flat].sum()) - 1) < 2e-3
    for s in scn.values():
        s.close()
