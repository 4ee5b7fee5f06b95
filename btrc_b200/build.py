"""In-tree build of libbtrc_b200.so (+ the `btrc` CLI) for sm_100a.

    python -m btrc_b200.build            # incremental
    python -m btrc_b200.build --force

Host sources are compiled with the system g++ (-ffp-contract=off: scene tables are bit-reproducible against the
oracle's independent loader), device sources with nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, and the
library is linked with the static CUDA runtime so that it can live beside torch's own runtime in one process.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbtrc_b200.so")
CLI = os.path.join(HERE, "btrc")

CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

HOST_SRCS = ["host/node.cpp", "host/mesh.cpp", "host/bvh.cpp", "host/scene.cpp", "host/renderer.cpp", "host/image_io.cpp", "host/comm.cpp", "capi.cpp"]
DEVICE_SRCS = ["device/kernels.cu", "device/lbvh.cu"]

HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-Wall", "-Wno-unknown-pragmas", f"-I{CUDA_HOME}/include"]
# -ftz / -prec-div / -prec-sqrt: the reference's own floating-point contract (fast_math => div.approx, sqrt.approx,
# flush-to-zero; log/exp/pow/atan2/acos stay full precision, so no --use_fast_math).  The triangle test and
# intersection_offset use explicit round-to-nearest intrinsics and are not affected.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", CXX,
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-ftz=true", "-prec-div=false", "-prec-sqrt=false"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _all_headers():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".h", ".cuh"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "btrc_b200.h"))
    return out


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get("BTRC_NVCC_EXTRA", "").split()  # experiment knobs, e.g. -DBTRC_SHADE_BLOCKS_HEAVY=3
    headers = _all_headers()
    jobs = []
    objs = []
    for src in HOST_SRCS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([CXX] + HOST_FLAGS + ["-c", s, "-o", o])
    for src in DEVICE_SRCS:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out.strip():
                print(out)
    if force or jobs or _newer(LIB, objs):
        _run([NVCC, "-shared", "-cudart", "static", "-ccbin", CXX, "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lz", "-lgomp", "-ldl"])
    cli_src = os.path.join(CSRC, "cli", "main.cpp")
    if force or _newer(CLI, [cli_src, LIB]):
        _run([CXX, "-std=c++17", "-O2", cli_src, "-o", CLI, f"-L{HERE}", "-lbtrc_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
