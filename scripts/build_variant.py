#!/usr/bin/env python
"""Full variant build of libbtrc_b200.so (host AND device objects recompiled with extra -D flags) into
btrc_b200/variants/libbtrc_<name>.so for A/B runs (`BTRC_B200_LIB=<path>`).

    python scripts/build_variant.py ctlpad -DBTRC_CTL_PAD=1
"""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from btrc_b200 import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
obj = os.path.join(B.OBJ, "variants", name)
os.makedirs(obj, exist_ok=True)
os.makedirs(os.path.join(B.HERE, "variants"), exist_ok=True)
jobs, objs = [], []
for src in B.HOST_SRCS:
    o = os.path.join(obj, src.replace("/", "_") + ".o")
    objs.append(o)
    jobs.append([B.CXX] + B.HOST_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", o])
for src in B.DEVICE_SRCS:
    o = os.path.join(obj, src.replace("/", "_") + ".o")
    objs.append(o)
    jobs.append([B.NVCC] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", o])
with cf.ThreadPoolExecutor(max_workers=8) as ex:
    list(ex.map(B._run, jobs))
lib = os.path.join(B.HERE, "variants", "libbtrc_%s.so" % name)
B._run([B.NVCC, "-shared", "-cudart", "static", "-ccbin", B.CXX, "-o", lib] + objs + ["-Xcompiler", "-fopenmp", "-lz", "-lgomp", "-ldl"])
print(lib)
