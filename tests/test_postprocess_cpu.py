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
    assert err.mean() < 3.0 and err.max() < 40  # quality-90 baseline JPEG
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
    """ImageDynamic::load for .jpg (the reference goes through stb_image): baseline files with 4:4:4 / 4:2:2 / 4:2:0 chroma, grayscale
    and restart intervals decode to within 3 eight-bit levels of libjpeg (PIL); progressive files are refused with a clear error."""
    from btrc_b200 import api

    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:67, 0:101]
    img = np.stack([127 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0), 30 + xx * 2.0, 200 - yy * 2.5], -1) + rng.normal(0, 6, (67, 101, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    cases = {"q90_420.jpg": dict(quality=90), "q75_444.jpg": dict(quality=75, subsampling=0), "q95_422.jpg": dict(quality=95, subsampling=1),
             "gray.jpg": dict(quality=85), "rst.jpg": dict(quality=80, restart_marker_blocks=5)}
    for name, kw in cases.items():
        p = str(tmp_path / name)
        Image.fromarray(img[..., 0] if name == "gray.jpg" else img).save(p, **kw)
        ref = np.asarray(Image.open(p))
        ref = ref if ref.ndim == 3 else ref[..., None]
        got = api.load_image(p)
        assert got.shape == ref.shape and got.dtype == np.uint8, name
        err = np.abs(got.astype(int) - ref.astype(int))
        assert err.max() <= 3 and err.mean() < 0.1, (name, err.max(), err.mean())
    p = str(tmp_path / "progressive.jpg")
    Image.fromarray(img).save(p, progressive=True)
    with pytest.raises(api.BtrcError, match="progressive"):
        api.load_image(p)
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
