#!/usr/bin/env python
"""State-pool size sweep on one workload: ms per render (CUDA events, profiling off) for several state_count values."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from btrc_b200 import api

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_disney_1080p_256spp"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pools = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24]
tmp = tempfile.TemporaryDirectory()
path = bench.make_scene(wl, tmp.name, spp=spp)
for pool in pools:
    s = api.Scene(path, state_count=pool)
    s.set_profiling(False)
    stream = torch.cuda.ExternalStream(s.stream_handle())
    W, H = s.info.width, s.info.height
    color = torch.zeros(H, W, 4, device="cuda")
    out = []
    for k in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        s.film_clear()
        st = s.render_range(0, W * H * spp)
        s.resolve(color)
        b.record(stream)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    ms = min(out[1:])
    print("pool %9d  ms %.2f  Msamples/s %.1f  iters %d  launches %d" % (pool, ms, W * H * spp / ms / 1e3, st["n_iter"], st["n_launches"]), flush=True)
    s.close()
    del s
