#!/usr/bin/env python
"""Mutation fuzzing of the texture readers (btrc_b200/csrc/host/image_io.cpp: PNG, JPEG baseline + progressive, Radiance HDR) under
AddressSanitizer + UBSan.  Texture files are outside input: a damaged one must end in the loader's error, never in an out-of-bounds
access, an undefined shift or a multi-gigabyte allocation from a corrupted header.

    python scripts/fuzz_image_io.py [mutations per seed file = 3000] [rng seed = 1]

Builds a host-only harness (g++ -fsanitize=address,undefined) in a temporary directory, writes the seed files (every decoder
branch: 16-bit / palette / 1-bit / RGBA PNGs with Adam7 and tRNS, baseline and progressive JPEGs with 4:4:4 / 4:2:2 / 4:2:0,
restart intervals, grayscale; one HDR), mutates them (byte flips, header bytes, truncation, deletions) and loads every mutant.
Exit code 0 = no sanitizer report.  tests/test_postprocess_cpu.py runs it with a small count."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MAIN = r'''
#include "image_io.h"
#include <cstdio>
#include <exception>
int main(int argc, char **argv) {
    int ok = 0, bad = 0;
    for (int i = 1; i < argc; ++i) {
        try { auto im = btrc::load_image(argv[i]); ok += im.width > 0; } catch (const std::exception &) { ++bad; }
    }
    std::printf("ok %d bad %d\n", ok, bad);
}
'''


def write_seeds(d, rng):
    from PIL import Image

    from tests.test_postprocess_cpu import _png_bytes

    yy, xx = np.mgrid[0:37, 0:45]
    img = np.stack([127 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0), 30 + xx * 4.0, 200 - yy * 4.5], -1) + rng.normal(0, 6, (37, 45, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    Image.fromarray(img).save(f"{d}/a.jpg", quality=80)
    Image.fromarray(img).save(f"{d}/b.jpg", quality=80, progressive=True)
    Image.fromarray(img).save(f"{d}/c.jpg", quality=60, progressive=True, subsampling=0, restart_marker_blocks=2)
    Image.fromarray(img[..., 0]).save(f"{d}/d.jpg", quality=80, progressive=True)
    Image.fromarray(img).save(f"{d}/e.jpg", quality=80, subsampling=1, restart_marker_blocks=3)
    u32 = np.uint32
    open(f"{d}/a.png", "wb").write(_png_bytes(rng.integers(0, 65536, (19, 23, 3)).astype(u32), 16, 2, True, trns=[0, 1, 0, 2, 0, 3]))
    open(f"{d}/b.png", "wb").write(_png_bytes(rng.integers(0, 4, (19, 23, 1)).astype(u32), 2, 3, True, plte=list(range(12)), trns=[1, 2]))
    open(f"{d}/c.png", "wb").write(_png_bytes(rng.integers(0, 256, (19, 23, 4)).astype(u32), 8, 6, False))
    open(f"{d}/d.png", "wb").write(_png_bytes(rng.integers(0, 2, (19, 23, 1)).astype(u32), 1, 0, True))
    w, h = 9, 5
    open(f"{d}/a.hdr", "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (h, w) + bytes(rng.integers(0, 256, w * h * 4, dtype=np.uint8)))


def mutate(data, rng):
    m = bytearray(data)
    mode = rng.integers(0, 4)
    if mode == 0:
        for _ in range(rng.integers(1, 4)):
            m[rng.integers(0, len(m))] = rng.integers(0, 256)
    elif mode == 1:
        m[rng.integers(0, min(len(m), 80))] = rng.integers(0, 256)  # header region
    elif mode == 2:
        m = m[:rng.integers(1, len(m))]
    else:
        i = rng.integers(0, len(m))
        del m[i:rng.integers(i, min(len(m), i + 40))]
    return m


def run(per_seed=3000, seed=1):
    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as d:
        open(f"{d}/main.cpp", "w").write(MAIN)
        host = os.path.join(ROOT, "btrc_b200", "csrc", "host")
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", host,
                        f"{d}/main.cpp", os.path.join(host, "image_io.cpp"), "-lz", "-o", f"{d}/fuzz"], check=True)
        os.makedirs(f"{d}/seeds")
        os.makedirs(f"{d}/mut")
        write_seeds(f"{d}/seeds", rng)
        files = []
        for name in sorted(os.listdir(f"{d}/seeds")):
            data = open(f"{d}/seeds/{name}", "rb").read()
            ext = name.split(".")[-1]
            files.append(f"{d}/seeds/{name}")
            for k in range(per_seed):
                p = f"{d}/mut/{name[0]}{k}.{ext}"
                open(p, "wb").write(mutate(data, rng))
                files.append(p)
        ok = bad = 0
        for i in range(0, len(files), 1000):
            r = subprocess.run([f"{d}/fuzz"] + files[i:i + 1000], capture_output=True, text=True)
            if r.returncode != 0 or "runtime error" in r.stderr or "AddressSanitizer" in r.stderr:
                sys.stderr.write(r.stderr[-4000:])
                return 1, ok, bad
            a, b = r.stdout.split()[1], r.stdout.split()[3]
            ok += int(a)
            bad += int(b)
        return 0, ok, bad


if __name__ == "__main__":
    rc, ok, bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 3000, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print(f"loaded {ok}, refused {bad}, sanitizer reports: {'yes' if rc else 'none'}")
    sys.exit(rc)
