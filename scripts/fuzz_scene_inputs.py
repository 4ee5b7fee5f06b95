#!/usr/bin/env python
"""Mutation fuzzing of the scene-file front end (btrc_b200/csrc/host/node.cpp: JSON with comments, $include, $reference, typed
values and transforms; mesh.cpp: OBJ loader and shading tables) under AddressSanitizer + UBSan - the companion of
scripts/fuzz_image_io.py.  Host only, no GPU.

    python scripts/fuzz_scene_inputs.py [mutations per seed file = 1500] [rng seed = 1]

Seeds: the material-test and medium scenes of btrc_b200/scenes.py, a UV-sphere OBJ, the Cornell box OBJ and a small OBJ with quads,
negative indices and missing attributes.  Every parsed tree is walked and every node is read as vec3 / transform / angle / scalar.
Exit code 0 = no sanitizer report."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MAIN = r'''
#include "mesh.h"
#include "node.h"
#include <cstdio>
#include <exception>
#include <string>
static void walk(const btrc::NodeP &n, int depth) {
    if (!n || depth > 64) return;
    try { n->as_vec3(); } catch (const std::exception &) {}
    try { n->as_transform3d(); } catch (const std::exception &) {}
    try { n->as_transform2d(); } catch (const std::exception &) {}
    try { n->as_degree(); } catch (const std::exception &) {}
    try { n->as_int(); n->as_float(); n->as_bool(); } catch (const std::exception &) {}
    for (auto &e : n->elements) walk(e, depth + 1);
    for (auto &c : n->children) walk(c.second, depth + 1);
}
int main(int argc, char **argv) {
    int ok = 0, bad = 0;
    for (int i = 1; i < argc; ++i) {
        std::string f = argv[i];
        try {
            if (f.size() > 4 && f.substr(f.size() - 4) == ".obj") { auto m = btrc::load_triangle_mesh(f, i & 1); ok += m.ntri >= 0; }
            else { auto n = btrc::parse_scene_file(f); walk(n, 0); ++ok; }
        } catch (const std::exception &) { ++bad; }
    }
    std::printf("ok %d bad %d\n", ok, bad);
}
'''

STRUCTURAL = list(b'{}[]",:$/-0123456789 \n')


def mutate(data, rng):
    m = bytearray(data)
    mode = rng.integers(0, 5)
    if mode == 0:
        for _ in range(rng.integers(1, 4)):
            m[rng.integers(0, len(m))] = rng.integers(0, 256)
    elif mode == 1:
        for _ in range(rng.integers(1, 4)):
            m[rng.integers(0, len(m))] = rng.choice(STRUCTURAL)
    elif mode == 2:
        m = m[:rng.integers(1, len(m))]
    elif mode == 3:
        i = rng.integers(0, len(m))
        del m[i:rng.integers(i, min(len(m), i + 40))]
    else:
        i = rng.integers(0, len(m))
        j = rng.integers(i, min(len(m), i + 60))
        m[i:i] = m[i:j] * int(rng.integers(1, 4))
    return m


def run(per_seed=1500, seed=1):
    from btrc_b200 import scenes

    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as d:
        open(f"{d}/main.cpp", "w").write(MAIN)
        host = os.path.join(ROOT, "btrc_b200", "csrc", "host")
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", host,
                        f"{d}/main.cpp", os.path.join(host, "mesh.cpp"), os.path.join(host, "node.cpp"), "-o", f"{d}/fuzz"], check=True)
        a = scenes.make_material_test(f"{d}/m", width=32, height=18, spp=1, rings=4, segments=6, tex_res=8, state_count=1000)
        b = scenes.make_medium_scene(f"{d}/v", width=32, height=18, spp=1, grid_res=4, state_count=1000)
        open(f"{d}/q.obj", "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nvt 1 0\nvt 1 1\nf 1/1/1 2/2/1 3/3/1 4//1\nf -4 -3 -2\nf 1//1 2//1 3//1\nv 2 0.5 0\nv 0.5 0.4 0.1\nf 1 2 5 3 6 4\nf 1 2 3 4 5 6 1 2\n")
        seeds = [a, b, os.path.join(f"{d}/m", sorted(f for f in os.listdir(f"{d}/m") if f.endswith(".obj"))[0]),
                 os.path.join(ROOT, "scenes", "assets", "cornell_box.obj"), f"{d}/q.obj"]
        os.makedirs(f"{d}/mut")
        files = list(seeds)
        for si, name in enumerate(seeds):
            data = open(name, "rb").read()
            for k in range(per_seed):
                p = f"{d}/mut/s{si}_{k}.{name.split('.')[-1]}"
                open(p, "wb").write(mutate(data, rng))
                files.append(p)
        ok = bad = 0
        for i in range(0, len(files), 500):
            r = subprocess.run([f"{d}/fuzz"] + files[i:i + 500], capture_output=True, text=True)
            if r.returncode != 0 or "runtime error" in r.stderr or "AddressSanitizer" in r.stderr:
                sys.stderr.write(r.stderr[-4000:])
                return 1, ok, bad
            ok += int(r.stdout.split()[1])
            bad += int(r.stdout.split()[3])
        return 0, ok, bad


if __name__ == "__main__":
    rc, ok, bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 1500, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print(f"loaded {ok}, refused {bad}, sanitizer reports: {'yes' if rc else 'none'}")
    sys.exit(rc)
