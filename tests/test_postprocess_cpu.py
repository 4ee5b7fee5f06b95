"""Post-processors on host images (cli/main.cpp:75-80; postprocess/save_to_image.cpp:30-62, tonemap.cpp:14-51; utils/image.cpp:16-114):
`save` to every format the reference writes (.png .jpg .hdr .exr) and the ACES tonemap -- no GPU involved."""
import json
import os
import struct

import numpy as np
import pytest
from PIL import Image

from tests.util import f32


def _read_exr(path):
    """Minimal reader for the uncompressed float scanline files tinyexr / the product write."""
    b = open(path, "rb").read()
    assert struct.unpack_from("<iI", b, 0) == (20000630, 2)
    p = 8
    attrs = {}
    while b[p] != 0:
        e = b.index(b"\0", p); name = b[p:e].decode(); p = e + 1
        e = b.index(b"\0", p); typ = b[p:e].decode(); p = e + 1
        size = struct.unpack_from("<i", b, p)[0]; p += 4
        attrs[name] = (typ, b[p:p + size]); p += size
    p += 1
    chans = []
    cl = attrs["channels"][1]
    q = 0
    while cl[q] != 0:
        e = cl.index(b"\0", q); chans.append(cl[q:e].decode()); q = e + 1
        assert struct.unpack_from("<i", cl, q)[0] == 2  # FLOAT
        q += 16
    assert attrs["compression"][1] == b"\0"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack_from("<%dQ" % h, b, p)
    img = np.zeros((h, w, len(chans)), f32)
    for y in range(h):
        yy, nbytes = struct.unpack_from("<ii", b, offs[y])
        assert yy == y and nbytes == 4 * w * len(chans)
        row = np.frombuffer(b, f32, w * len(chans), offs[y] + 8).reshape(len(chans), w)
        img[y] = row.T
    return chans, img


def test_save_all_formats_and_tonemap(built, tmp_path):
    from btrc_b200 import api, scenes

    W, H = 40, 24
    path = scenes.make_cornell(str(tmp_path), width=W, height=H, spp=1, state_count=1024)
    root = json.load(open(path))
    root["post_processors"] = [{"type": "save", "color": "$env{scene_directory}/raw.exr"},
                               {"type": "save", "color": "$env{scene_directory}/raw.hdr"},
                               {"type": "tonemap", "exposure": 1.5},
                               {"type": "save", "color": "$env{scene_directory}/out.png", "gamma": 2.2},
                               {"type": "save", "color": "$env{scene_directory}/out.jpg", "gamma": 2.2}]
    json.dump(root, open(path, "w"))
    s = api.Scene(path)
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:H, 0:W]
    color = np.zeros((H, W, 4), f32)
    color[..., 0] = (xx / W) * 3.0
    color[..., 1] = (yy / H) * 0.8
    color[..., 2] = 0.25 + 0.1 * rng.random((H, W))
    color[..., 3] = 1.0
    hdr_in = color.copy()
    s.run_post_processors(color)

    chans, exr = _read_exr(str(tmp_path / "raw.exr"))
    assert chans == ["B", "G", "R"]
    assert np.array_equal(exr[..., ::-1], hdr_in[..., :3])  # float data, bit exact

    b = open(tmp_path / "raw.hdr", "rb").read()
    head_end = b.index(b"\n\n") + 2
    line_end = b.index(b"\n", head_end)
    assert b[head_end:line_end] == b"-Y %d +X %d" % (H, W)
    rgbe = np.frombuffer(b, np.uint8, W * H * 4, line_end + 1).reshape(H, W, 4).astype(np.float64)
    dec = rgbe[..., :3] * (2.0 ** (rgbe[..., 3:4] - 136.0))
    # shared exponent: every component is quantised to 1/256 of the pixel's largest one (truncated)
    assert (np.abs(dec - hdr_in[..., :3]) <= hdr_in[..., :3].max(axis=-1, keepdims=True) / 128 + 1e-6).all()

    # tonemap (in place on the caller's image): ACES fit of exposure * colour (tonemap.cpp:14-22)
    x = 1.5 * hdr_in[..., :3].astype(np.float64)
    aces = np.clip((x * (2.51 * x + 0.03)) / (x * (2.43 * x + 0.59) + 0.14), 0, 1)
    assert np.allclose(color[..., :3], aces, rtol=1e-5, atol=1e-6)

    want8 = np.minimum((aces ** (1 / 2.2) * 256).astype(np.int64), 255).astype(np.uint8)
    png = np.asarray(Image.open(tmp_path / "out.png"))
    assert png.shape == (H, W, 4)
    assert np.abs(png[..., :3].astype(int) - want8.astype(int)).max() <= 1
    jpg = np.asarray(Image.open(tmp_path / "out.jpg").convert("RGB"))
    assert jpg.shape == (H, W, 3)
    err = np.abs(jpg.astype(int) - want8.astype(int))
    assert err.mean() < 5.0 and err.max() < 60  # quality-90 baseline JPEG with 2x2 chroma subsampling, as stb_image_write writes it (pinned in test_reference_host_pins_cpu.py)
    s.close()


def test_unknown_format_is_an_error(built, tmp_path):
    from btrc_b200 import api, scenes

    path = scenes.make_cornell(str(tmp_path), width=8, height=8, spp=1, state_count=64)
    root = json.load(open(path))
    root["post_processors"] = [{"type": "save", "color": "$env{scene_directory}/x.tiff"}]
    json.dump(root, open(path, "w"))
    s = api.Scene(path)
    with pytest.raises(api.BtrcError, match="unknown image format"):
        s.run_post_processors(np.zeros((8, 8, 4), f32))
    s.close()


def test_jpeg_decoder_against_libjpeg(built, tmp_path):
    """ImageDynamic::load for .jpg (the reference goes through stb_image): baseline and progressive files with 4:4:4 / 4:2:2 / 4:2:0
    chroma, grayscale and restart intervals decode close to libjpeg (PIL; the byte-exact pin is against stb_image itself, in
    tests/test_reference_host_pins_cpu.py, which also runs without the reference tree's prebuilt library being rebuilt); arithmetic-coded
    files (which stb_image does not read either) are refused with a clear error."""
    from btrc_b200 import api

    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:67, 0:101]
    img = np.stack([127 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0), 30 + xx * 2.0, 200 - yy * 2.5], -1) + rng.normal(0, 6, (67, 101, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    cases = {"q90_420.jpg": dict(quality=90), "q75_444.jpg": dict(quality=75, subsampling=0), "q95_422.jpg": dict(quality=95, subsampling=1),
             "gray.jpg": dict(quality=85), "rst.jpg": dict(quality=80, restart_marker_blocks=5),
             # progressive (SOF2): DC first + refinement, AC bands with end-of-band runs, AC refinement, single-component scans
             "prog_420.jpg": dict(quality=90, progressive=True), "prog_444.jpg": dict(quality=97, progressive=True, subsampling=0),
             "prog_422.jpg": dict(quality=60, progressive=True, subsampling=1), "prog_gray.jpg": dict(quality=85, progressive=True),
             "prog_rst.jpg": dict(quality=80, progressive=True, restart_marker_blocks=3), "prog_lowq.jpg": dict(quality=10, progressive=True)}
    for name, kw in cases.items():
        p = str(tmp_path / name)
        Image.fromarray(img[..., 0] if "gray" in name else img).save(p, **kw)
        ref = np.asarray(Image.open(p))
        ref = ref if ref.ndim == 3 else ref[..., None]
        got = api.load_image(p)
        assert got.shape == ref.shape and got.dtype == np.uint8, name
        err = np.abs(got.astype(int) - ref.astype(int))
        # the decoder reproduces stb_image, the reference's decoder, byte for byte (tests/test_reference_host_pins_cpu.py); libjpeg differs from
        # stb_image in its IDCT rounding and in the chroma filter at the right edge (stb's last 4:2:2 pair): a sanity bound, not the parity bar
        assert err.max() <= 12 and err.mean() < 0.5, (name, err.max(), err.mean())
        if "prog" in name:
            assert b"\xff\xc2" in open(p, "rb").read()  # really a progressive frame
    # a file whose frame header says SOF9 (arithmetic coding): refused, as by stb_image
    data = open(tmp_path / "q90_420.jpg", "rb").read().replace(b"\xff\xc0", b"\xff\xc9", 1)
    open(tmp_path / "arith.jpg", "wb").write(data)
    with pytest.raises(api.BtrcError, match="arithmetic"):
        api.load_image(str(tmp_path / "arith.jpg"))
    # truncated entropy data: an error or a partial picture, never a crash
    open(tmp_path / "cut.jpg", "wb").write(open(tmp_path / "prog_420.jpg", "rb").read()[:900])
    try:
        assert api.load_image(str(tmp_path / "cut.jpg")).shape == (67, 101, 3)
    except api.BtrcError:
        pass
    # the product's own writer -> its own reader (and PIL's): what `save` writes can be used as a texture
    from btrc_b200 import scenes
    path = scenes.make_cornell(str(tmp_path), width=101, height=67, spp=1, state_count=64)
    root = json.load(open(path))
    root["post_processors"] = [{"type": "save", "color": "$env{scene_directory}/own.jpg"}]
    json.dump(root, open(path, "w"))
    s = api.Scene(path)
    color = np.concatenate([img.astype(f32) / f32(256) + f32(0.5 / 256), np.ones((67, 101, 1), f32)], -1)
    s.run_post_processors(np.ascontiguousarray(color))
    own = api.load_image(str(tmp_path / "own.jpg"))
    pil = np.asarray(Image.open(tmp_path / "own.jpg"))
    assert np.abs(own.astype(int) - pil.astype(int)).max() <= 3
    assert np.abs(own.astype(int) - img.astype(int)).mean() < 6.0   # quality-90 quantisation of an image with sigma-6 noise
    s.close()


def test_png_and_hdr_load(built, tmp_path):
    from btrc_b200 import api

    rng = np.random.default_rng(4)
    rgba = rng.integers(0, 256, (13, 17, 4), dtype=np.uint8)
    Image.fromarray(rgba, "RGBA").save(tmp_path / "a.png")
    assert np.array_equal(api.load_image(str(tmp_path / "a.png")), rgba)
    Image.fromarray(rgba[..., 0], "L").save(tmp_path / "g.png")
    assert np.array_equal(api.load_image(str(tmp_path / "g.png"))[..., 0], rgba[..., 0])
    with pytest.raises(api.BtrcError, match="unsupported image format"):
        api.load_image(str(tmp_path / "x.exr"))


def _png_bytes(samples, depth, ctype, interlace=False, plte=None, trns=None):
    """A PNG file from an (H, W, C) array of raw sample values: any bit depth / colour type, optional Adam7, the five filter
    types in rotation (test-side writer: PIL writes neither interlaced files nor 16-bit RGB)."""
    import struct
    import zlib

    h, w, ch = samples.shape

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    def pack_row(row):  # (pw, ch) raw samples -> bytes
        flat = row.reshape(-1).astype(np.uint32)
        if depth == 16:
            return flat.astype(">u2").tobytes()
        if depth == 8:
            return flat.astype(np.uint8).tobytes()
        bits = np.zeros((flat.size * depth + 7) // 8 * 8, np.uint8)
        for b in range(depth):
            bits[b:flat.size * depth:depth] = (flat >> (depth - 1 - b)) & 1
        return np.packbits(bits).tobytes()

    def paeth(a, b, c):
        pp = a + b - c
        pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
        return a if pa <= pb and pa <= pc else (b if pb <= pc else c)

    bpp = max(1, depth * ch // 8)
    raw = bytearray()
    nrow = 0
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    for x0, y0, dx, dy in passes:
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        prev = None
        for r in range(sub.shape[0]):
            cur = pack_row(sub[r])
            up = prev if prev is not None else bytes(len(cur))
            ft = nrow % 5
            nrow += 1
            out = bytearray([ft])
            for i, v in enumerate(cur):
                a = cur[i - bpp] if i >= bpp else 0
                b = up[i]
                c = up[i - bpp] if i >= bpp else 0
                pred = (0, a, b, (a + b) // 2, paeth(a, b, c))[ft]
                out.append((v - pred) & 255)
            raw += out
            prev = cur
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if plte is not None:
        data += chunk(b"PLTE", bytes(plte))
    if trns is not None:
        data += chunk(b"tRNS", bytes(trns))
    z = zlib.compress(bytes(raw))
    data += chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    return data


def test_png_every_variant_stb_reads(built, tmp_path):
    """ImageDynamic::load for .png goes through stb_image's 8-bit interface (utils/image.cpp:116-147): every colour type and bit
    depth, Adam7 interlacing, palettes and tRNS colour keys, with stb_image's conversions (high byte of 16-bit samples, grey
    scaled by 255 / (2^depth - 1), palette -> RGB / RGBA, colour key -> alpha 0 / 255).  Files are also opened with PIL to make
    sure the test-side writer produced valid PNGs."""
    from btrc_b200 import api

    rng = np.random.default_rng(8)
    H, W = 19, 23   # not multiples of 8: ragged Adam7 passes and ragged packed rows
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    n = 0
    for interlace in (False, True):
        for ctype, ch, depths in ((0, 1, (1, 2, 4, 8, 16)), (2, 3, (8, 16)), (3, 1, (1, 2, 4, 8)), (4, 2, (8, 16)), (6, 4, (8, 16))):
            for depth in depths:
                hi = min(1 << depth, 16) if ctype == 3 else 1 << depth
                smp = rng.integers(0, hi, (H, W, ch)).astype(np.uint32)
                name = str(tmp_path / f"t{ctype}_d{depth}_i{int(interlace)}.png")
                open(name, "wb").write(_png_bytes(smp, depth, ctype, interlace, plte=pal.reshape(-1) if ctype == 3 else None))
                if ctype == 3:
                    want = pal[smp[..., 0]]
                elif depth == 16:
                    want = (smp >> 8).astype(np.uint8)
                else:
                    want = (smp * (255 // ((1 << depth) - 1))).astype(np.uint8)
                if ctype == 4:   # grey + alpha: ImageDynamic::load has no two-channel image type (utils/image.cpp:333-334)
                    with pytest.raises(api.BtrcError, match="unsupported channels: 2"):
                        api.load_image(name)
                    n += 1
                    continue
                got = api.load_image(name)
                assert got.shape == want.shape and np.array_equal(got, want), name
                pil = Image.open(name)
                pil.load()
                assert pil.size == (W, H)
                if depth <= 8 and ctype in (2, 3, 6):
                    assert np.array_equal(np.asarray(pil.convert("RGBA" if ctype == 6 else "RGB")), want), name  # PIL agrees where it has no 16-bit modes in the way
                n += 1
    assert n == 30
    # tRNS: palette alpha, grey and RGB colour keys (8- and 16-bit)
    idx = rng.integers(0, 16, (H, W, 1)).astype(np.uint32)
    alpha = rng.integers(0, 256, 10, dtype=np.uint8)  # shorter than the palette: the rest is opaque
    open(tmp_path / "pt.png", "wb").write(_png_bytes(idx, 4, 3, True, plte=pal.reshape(-1), trns=alpha))
    got = api.load_image(str(tmp_path / "pt.png"))
    assert got.shape == (H, W, 4) and np.array_equal(got[..., :3], pal[idx[..., 0]])
    assert np.array_equal(got[..., 3], np.concatenate([alpha, np.full(6, 255, np.uint8)])[idx[..., 0]])
    assert np.array_equal(np.asarray(Image.open(tmp_path / "pt.png").convert("RGBA")), got)
    g = rng.integers(0, 4, (H, W, 1)).astype(np.uint32)
    open(tmp_path / "gk.png", "wb").write(_png_bytes(g, 2, 0, False, trns=[0, 2]))
    with pytest.raises(api.BtrcError, match="unsupported channels: 2"):   # a colour-keyed grey image becomes grey + alpha in stb_image
        api.load_image(str(tmp_path / "gk.png"))
    c16 = rng.integers(0, 3, (H, W, 3)).astype(np.uint32) * 0x0101 + 0x1000
    open(tmp_path / "ck.png", "wb").write(_png_bytes(c16, 16, 2, True, trns=[0x11, 0x01, 0x10, 0x00, 0x12, 0x02]))
    got = api.load_image(str(tmp_path / "ck.png"))
    key = (c16[..., 0] == 0x1101) & (c16[..., 1] == 0x1000) & (c16[..., 2] == 0x1202)
    assert key.any() and got.shape == (H, W, 4)
    assert np.array_equal(got[..., :3], (c16 >> 8).astype(np.uint8)) and np.array_equal(got[..., 3], np.where(key, 0, 255))
    # malformed files: errors, not crashes
    good = _png_bytes(rng.integers(0, 256, (H, W, 3)).astype(np.uint32), 8, 2)
    for label, bad in (("cut", good[:60]), ("depth", good[:24] + b"\x03" + good[25:]), ("sig", b"\x00" + good[1:])):
        open(tmp_path / f"{label}.png", "wb").write(bad)
        with pytest.raises(api.BtrcError, match="failed to load image"):
            api.load_image(str(tmp_path / f"{label}.png"))


def test_texture_readers_survive_damaged_files(built):
    """Mutation fuzzing of the PNG / JPEG / HDR readers under AddressSanitizer + UBSan (scripts/fuzz_image_io.py; 30,000 mutants per run
    there, 2,500 here): damaged texture files end in the loader's error - no out-of-bounds access, no undefined shift, no allocation
    sized by a corrupted header."""
    import importlib.util
    import shutil

    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    spec = importlib.util.spec_from_file_location("fuzz_image_io", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "fuzz_image_io.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rc, ok, bad = mod.run(per_seed=250, seed=5)
    assert rc == 0
    assert ok >= 10 and bad > 100   # the untouched seed files load, most mutants are refused
