#!/usr/bin/env python
"""Where does a bench step's time go?  Times single steps of the C2 workload under different brackets:
plain, with the 256 MB L2 flush, with / without stage profiling events, with / without the nvidia-smi sampler."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from btrc_b200 import api

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_disney_1080p_256spp"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tmp = tempfile.TemporaryDirectory()
path = bench.make_scene(wl, tmp.name, spp=spp)
s = api.Scene(path, state_count=2097152)
stream = torch.cuda.ExternalStream(s.stream_handle())
W, H = s.info.width, s.info.height
color = torch.zeros(H, W, 4, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(tag, n=4, do_flush=False, prof=False, smi=False):
    s.set_profiling(prof)
    sm = bench.ClockSampler(0)
    if smi:
        sm.start()
    out = []
    for k in range(n):
        if do_flush:
            flush.fill_(k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(stream)
        s.film_clear()
        st = s.render_range(0, W * H * spp)
        s.resolve(color)
        b.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        out.append((a.elapsed_time(b), wall, st["render_ms"], sum(st["stage_ms"].values())))
    if smi:
        sm.stop()
    print(tag, " | ".join("ev %.1f wall %.1f render %.1f stages %.1f" % o for o in out), flush=True)


run("warm        ")
run("plain       ")
run("flush       ", do_flush=True)
run("prof        ", prof=True)
run("prof+flush  ", prof=True, do_flush=True)
run("smi         ", smi=True)
run("smi+prof+fl ", smi=True, prof=True, do_flush=True)
run("plain again ")
