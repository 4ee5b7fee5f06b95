#!/usr/bin/env python
"""Benchmark of the wavefront render path (BASELINE.json metric: Msamples/s and Mrays/s per scene; % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference ...                   # CPU arm: the oracle restatement on all host cores

A "step" is one pass of the hot path over one batch of synthetic input: one full render of the workload scene
(film clear -> generate / closest-hit / shade / shadow waves until every camera path has ended -> [NCCL sum of the
films] -> resolve).  Rank r renders sample indices [r*spp, (r+1)*spp) of every pixel (sample-stream partition),
so per-GPU work is fixed as N grows ("weak").  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[0]: Cornell box, diffuse + one area light, 512x512, 16 spp
    "c1_cornell_512_16spp": dict(kind="cornell", width=512, height=512, spp=16),
    # BASELINE.json configs[1] (the configuration the metric is quoted on): Disney material test with image textures, 1920x1080, 256 spp
    "c2_disney_1080p_256spp": dict(kind="materials", width=1920, height=1080, spp=256),
    # BASELINE.json configs[2]: heterogeneous grid-volume medium with NEE, 1920x1080 (spp not stated => reference default 128)
    "c3_medium_1080p_128spp": dict(kind="medium", width=1920, height=1080, spp=128),
    # BASELINE.json configs[3]: synthetic 10M-triangle instanced mesh scene (250 x 40,000) with 64 area lights, max depth 16
    "c4_instanced_10m_1080p_128spp": dict(kind="instanced", width=1920, height=1080, spp=128),
    # BASELINE.json configs[4]: the C2 scene at 4K, 1024 spp in total, split over the ranks (strong scaling: --gpus 2/4/8)
    "c5_disney_4k_1024spp": dict(kind="materials", width=3840, height=2160, spp=1024, strong=True),
}
DEFAULT_WORKLOAD = "c2_disney_1080p_256spp"


def make_scene(workload, out_dir, spp=None, tile=False):
    from btrc_b200 import scenes

    w = WORKLOADS[workload]
    os.makedirs(out_dir, exist_ok=True)
    make = {"cornell": scenes.make_cornell, "materials": scenes.make_material_test, "medium": scenes.make_medium_scene,
            "instanced": scenes.make_instanced_scene}[w["kind"]]
    path = make(out_dir, width=w["width"], height=w["height"], spp=spp or w["spp"])
    if tile:  # Tile-mode path indexing (renderer "tile": true, generate.cpp:56-60) for the image-tile partition
        root = json.load(open(path))
        root["renderer"]["tile"] = True
        json.dump(root, open(path, "w"), indent=1)
    return path


def workload_config(workload):
    """The `config` object both arms print (same keys, so the driver can tell they ran the same workload)."""
    w = WORKLOADS[workload]
    return {"workload": workload, "width": w["width"], "height": w["height"], "spp": w["spp"]}


def scaling_of(args):
    return "strong" if WORKLOADS[args.workload].get("strong") else "weak"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe) through NVML in a
    background thread (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; an external
    `nvidia-smi -lms 100` loop was measured to stall kernel launches for 100+ ms at a time -- profiles/r1h_diag_step.log)."""

    PERIOD_S = 0.25

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.thread = None
        self._stop = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(self.max_sm)
                if get_reasons is not None:
                    bits = int(get_reasons(self.h))
                    for name, attr, default in names:
                        if bits & int(getattr(nv, attr, default)):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def reset(self):
        """Drops the samples taken so far (warm-up): only the timed region counts."""
        self.sm, self.mx, self.reasons = [], [], set()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = sorted(self.sm)[len(self.sm) // 2:]  # samples under load are the upper half (the GPU idles between steps)
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(self.mx)), "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "NVML, %.2f s period" % self.PERIOD_S}


# ------------------------------------------------------------------------------------------------ algorithmic bytes
def bvh_alg_per_ray(scene_path, sample_paths=16384):
    """lib/bvh binary-BVH statistics: mean traversal steps and primitive tests per closest-hit ray, measured
    by oracle/_ref (the reference's vendored SingleRayTraverser, sweep-SAH) on the ray batches of a bounded path sample."""
    import ctypes

    from oracle import front
    from oracle.front import _ptr

    R = front.ref_lib()
    if R is None:
        return None
    o = front.OracleScene(scene_path)
    verts, _ = o.triangles()
    h = R.ref_bvh_build(verts.shape[0], _ptr(np.ascontiguousarray(verts), ctypes.c_float), 0)
    # a sample of camera paths spread over the whole image (64 chunks of consecutive path indices), not its first rows
    npix = o.width * o.height
    chunks = [o.generate((k * npix) // 64 + (k * 37) % max(o.width - sample_paths // 64, 1), sample_paths // 64) for k in range(64)]
    st = {k: np.ascontiguousarray(np.concatenate([c[k] for c in chunks])) for k in chunks[0]}
    stats = np.zeros(2, np.uint64)
    n_rays = 0
    for _ in range(64):
        n = st["o_med"].shape[0]
        if n == 0:
            break
        prim = np.zeros(n, np.int32)
        tuv = np.zeros((n, 3), np.float32)
        R.ref_bvh_closest(h, n, _ptr(st["o_med"], ctypes.c_float), _ptr(st["d_t1"], ctypes.c_float), _ptr(prim, ctypes.c_int32),
                          _ptr(tuv, ctypes.c_float), _ptr(stats, ctypes.c_uint64))
        n_rays += n
        flag, detail = o.closest_hit(st["o_med"], st["d_t1"])
        parts = []
        if o.has_medium():
            st, flag, mo = o.medium(st, flag, detail)
            parts.append({k: v[(mo["status"] & 1) != 0] for k, v in mo["next"].items()})
        sh = o.shade(st, flag, detail)
        parts.append({k: v[(sh["status"] & 1) != 0] for k, v in sh["next"].items()})
        st = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in parts[0]}
    R.ref_bvh_free(h)
    return {"steps_per_ray": float(stats[0]) / n_rays, "tests_per_ray": float(stats[1]) / n_rays, "rays": n_rays}


def stage_bytes(stats, bvh, has_medium=False):
    """Algorithmic bytes of one render from the reference SoA layout."""
    bvh_per_ray = (bvh["steps_per_ray"] * 64 + bvh["tests_per_ray"] * 48) if bvh else 0.0
    n_closest, n_hit = stats["n_closest"], stats["n_hit"]
    trace = n_closest * 32 + n_hit * 20 + (n_closest - n_hit) * 4 + n_closest * bvh_per_ray
    shade = n_closest * 108 + n_hit * (16 + 96) + stats["n_next"] * 104 + stats["n_shadow"] * 56 + (stats["n_splat"] - stats["n_shadow_miss"]) * 12
    shadow = stats["n_shadow"] * (32 + bvh_per_ray) + stats["n_shadow_miss"] * (24 + 12)
    generate = stats["n_gen"] * 108
    # medium kernel (if run): 108 R + (16 + 16 on hit) + 32 + 16 + 4 W per state; its scatter outputs (104 / 56 W) are counted in shade's
    # n_next / n_shadow terms above, which are whole-iteration totals of both kernels
    medium = (n_closest * (108 + 52) + n_hit * 32) if has_medium else 0
    return {"generate": generate, "trace": trace, "shade": shade, "shadow": shadow, "medium": medium}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm
class CpuArm:
    """The reference's wavefront renderer on the host cores.  kind "reference": oracle/_ref/libbtrcrefcpu.so -- Btrc's own
    scene loading, kernels (through CUJ's LLVM-free C++ generator) and host loop compiled from /root/reference where it lies,
    lib/bvh standing in for OptiX (oracle/refcpu/refcpu.cpp); the kernels of a scene are compiled with g++ when the scene is
    opened (the reference compiles its kernels per scene too, wavefront.cpp:141-179).  kind "port": the oracle restatement
    (oracle/liborc.so), used when that library or the host compiler is not there."""

    def __init__(self, scene_path):
        from oracle import front

        self.kind, self.note = "port", ""
        self.ref = None
        try:
            from oracle import refcpu

            if refcpu.available():
                os.environ.pop("OMP_NUM_THREADS", None)  # torchrun exports OMP_NUM_THREADS=1 to every rank
                t0 = time.perf_counter()
                self.ref = refcpu.ReferenceScene(scene_path)
                self.kind = "reference"
                self.note = "oracle/_ref/libbtrcrefcpu.so (reference sources + CUJ C++ back end, g++ -O1; kernels compiled in %.0f s, untimed)" % (time.perf_counter() - t0)
        except Exception as e:  # no host compiler on this box, or a scene feature the host build lacks
            self.note = "reference host build unavailable (%s); " % str(e).strip().splitlines()[-1][:120]
        if self.ref is None:
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
            front.lib().orc_use_all_cores()
            self.orc = front.OracleScene(scene_path)
            self.note += "oracle/liborc.so (restatement), OpenMP"
        self.cores = os.cpu_count() if self.ref is not None else front.lib().orc_num_threads()

    def render(self, spp):
        """Renders `spp` samples per pixel; returns (seconds, rays)."""
        t0 = time.perf_counter()
        if self.ref is not None:
            self.ref.set_spp(spp)
            r = self.ref.render()
        else:
            r = self.orc.render(0, spp)
        return time.perf_counter() - t0, r["stats"]["n_closest"] + r["stats"]["n_shadow"]


def run_reference(args):
    """Times the reference's own CPU execution of the path (CpuArm) on all host cores, each step a bounded sample of the
    workload (as many of its spp as fit the time budget)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    with tempfile.TemporaryDirectory() as d:
        path = make_scene(args.workload, d)
        arm = CpuArm(path)
        t1 = 0.0
        for _ in range(max(args.warmup, 1)):
            t1 = arm.render(1)[0]
        # each step = a bounded sample of the workload: as many spp as fit in ~60 s / steps of CPU time
        sample_spp = int(max(1, min(w["spp"], 60.0 / max(args.steps, 1) / max(t1, 1e-3))))
        dt, rays = 0.0, 0
        for k in range(args.steps):
            a, b = arm.render(sample_spp)
            dt += a
            rays += b
    paths = w["width"] * w["height"] * sample_spp * args.steps
    value = paths / dt / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "mrays_per_s": rays / dt / 1e6,
        "config": workload_config(args.workload),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": f"{sample_spp} of {w['spp']} spp per step ({w['width']}x{w['height']}); " + arm.note},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from btrc_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = WORKLOADS[args.workload]
    W, H, spp = w["width"], w["height"], (args.spp or w["spp"])
    strong = bool(w.get("strong"))
    if strong:  # total work fixed: every rank renders spp / world samples per pixel
        if spp % world:
            raise SystemExit(f"bench.py: {spp} spp do not split evenly over {world} ranks")
        spp //= world
    npix = W * H
    tmp = tempfile.TemporaryDirectory()
    tiles = args.partition == "tiles"
    # samples partition: the scene holds spp * world samples per pixel and every rank renders `spp` of them;
    # tiles partition: every rank renders all `spp * world` samples of its band of pixels (same number of paths per rank)
    path = make_scene(args.workload, tmp.name, spp=spp * world, tile=tiles)
    s = api.Scene(path, state_count=args.state_count)
    s.set_profiling(False)  # the timed steps carry no per-stage events; the stage breakdown comes from a separate profiled pass
    stream = torch.cuda.ExternalStream(s.stream_handle())
    from btrc_b200 import partition

    begin, end = (partition.tile_path_range if tiles else partition.path_range)(rank, world, spp * world, npix)

    rad = torch.zeros(H, W, 4, device="cuda")
    wgt = torch.zeros(H, W, device="cuda")
    color_dev = torch.zeros(H, W, 4, device="cuda")
    color_host = torch.zeros(H, W, 4).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """One device-resident step: returns the render statistics of this rank."""
        s.film_clear()
        st = s.render_range(begin, end)
        if world > 1:  # the one collective of the path: sum of the film buffers (radiance float4 + weight), then resolve on rank 0
            s.film_export(rad, wgt)
            partition.reduce_films([rad, wgt], 0)
            if rank == 0:
                s.film_import(rad, wgt)
        if rank == 0:
            s.resolve(color_dev)
        return st

    def step_e2e():
        """The same step through the public call with HOST buffers: result image read back to pinned host memory."""
        if world == 1:
            _, _, _, st = s.render(color_host.numpy())
            return st
        st = step()
        if rank == 0:
            color_host.copy_(color_dev, non_blocking=False)
        return st

    # NVML is initialised before the warm-up steps (its start-up takes driver locks that stall launches for ~100 ms)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    if rank == 0:
        sampler.reset()

    # ---- timed region: K device-resident steps, CUDA events on the renderer's stream, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    agg = None
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        st = step()
        ev[k][1].record(stream)
        if agg is None:
            agg = {key: 0 for key in st if key != "stage_ms"}
            agg["stage_ms"] = {key: 0.0 for key in st["stage_ms"]}
        for key, v in st.items():
            if key == "stage_ms":
                for kk, vv in v.items():
                    agg["stage_ms"][kk] += vv
            else:
                agg[key] += v
    barrier()
    wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # ---- stage breakdown: the same steps again with one CUDA event between consecutive kernels (not part of `value`)
    s.set_profiling(True)
    prof_steps = min(args.steps, 3)
    pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(prof_steps)]
    stage_ms = None
    prof_iters = 0
    for k in range(prof_steps):
        flush.fill_(k & 0xff)
        torch.cuda.synchronize()
        pev[k][0].record(stream)
        st = step()
        pev[k][1].record(stream)
        torch.cuda.synchronize()
        if stage_ms is None:
            stage_ms = {key: 0.0 for key in st["stage_ms"]}
        for kk, vv in st["stage_ms"].items():
            stage_ms[kk] += vv
        prof_iters += st["n_iter"]
    prof_ms = sum(a.elapsed_time(b) for a, b in pev)
    s.set_profiling(False)

    # ---- end to end through the public API with host buffers
    barrier()
    e2e_t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - e2e_t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # whole-job event counts
    counts = torch.tensor([agg[k] for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat", "n_iter", "n_launches")],
                          device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(counts)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_gen, n_closest, n_hit, n_next, n_shadow, n_shadow_miss, n_splat, n_iter, n_launch = [float(x) for x in counts.tolist()]
    secs = dev_ms / 1e3
    value = n_gen / secs / 1e6
    mrays = (n_closest + n_shadow) / secs / 1e6

    # ---- roofline of the dominant kernel (rank 0's stage times; algorithmic bytes per SURVEY 8d)
    bvh = None if args.no_cpu_baseline else bvh_alg_per_ray(path)
    peak, peak_src = measured_peak()
    sb = stage_bytes({k: agg[k] for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat")}, bvh,
                     has_medium=bool(s.info.has_medium))
    iters = max(prof_iters, 1)
    stages = {}
    sb["bookkeeping"] = 0
    scale = prof_steps / args.steps  # the byte totals cover K timed steps, the stage times `prof_steps` identical ones
    for name in ("generate", "trace", "medium", "shade", "shadow", "bookkeeping"):
        ms = stage_ms[name]
        if name == "medium" and not s.info.has_medium:
            continue
        nbytes = sb[name] * scale
        gbs = nbytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        stages[name] = {"ms_total": ms, "launches": iters, "ms_per_launch": ms / iters, "alg_bytes_per_launch": nbytes / iters,
                        "achieved_gbs": gbs, "frac": gbs / peak, "share_of_step": ms / prof_ms}
    dom = max((k for k in stages if k != "bookkeeping"), key=lambda k: stages[k]["ms_total"])
    # DRAM traffic of the dominant kernel: bytes per ray / state from the committed ncu --set full capture of this workload
    # (profiles/traffic_<workload>.json, written by scripts/ncu_summary.py traffic), scaled to this run's units per launch
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(dom)
        if t and t.get("dram_bytes_per_unit"):
            units = {"trace": agg["n_closest"], "shadow": agg["n_shadow"], "shade": agg["n_hit"]}.get(dom)
            if units:
                traffic = t["dram_bytes_per_unit"] * units * scale / iters
    roofline = {"kernel": dom + "_kernel",
                "bound": "hbm", "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": traffic, "traffic_source": ("profiles/traffic_%s.json (ncu dram__bytes_read+write per ray x rays per launch)" % args.workload) if traffic else None,
                "peak_source": peak_src, "alg_bytes_per_launch": stages[dom]["alg_bytes_per_launch"],
                "ms_per_launch": stages[dom]["ms_per_launch"], "bvh_alg": bvh}

    # ---- CPU baseline: the reference's own CPU execution of the path (CpuArm) on the host cores of this box, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        arm = CpuArm(make_scene(args.workload, os.path.join(tmp.name, "cpu")))
        t1, _ = arm.render(1)  # calibration pass: 1 spp
        cpu_spp = int(max(1, min(spp, 15.0 / max(t1, 1e-3))))  # ~15 s of CPU work
        cpu_s, _ = arm.render(cpu_spp)
        cpu = {"value": npix * cpu_spp / cpu_s / 1e6, "unit": "Msamples/s", "cores": arm.cores, "kind": arm.kind,
               "sample": f"{cpu_spp} of {spp} spp of the workload ({W}x{H}), {cpu_s:.1f} s on all host cores; " + arm.note}

    e2e_value = n_gen / e2e_s / 1e6
    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "mrays_per_s": mrays,
        "config": {"workload": args.workload, "width": W, "height": H, "spp_per_gpu": spp, "state_count": s.info.state_count,
                   "partition": (("image tiles (row bands of pixels, Tile-mode indexing) per rank" if tiles else "sample index ranges per rank") +
                                 "; NCCL sum-reduce of radiance+weight films to rank 0; resolve on rank 0") if world > 1 else "single GPU",
                   "timing": "CUDA events on the renderer's stream per step, max over ranks; 256 MB L2 flush (untimed) between steps; "
                             "roofline_stages from a separate pass of %d identical steps with one CUDA event between kernels (%.1f ms/step)" % (prof_steps, prof_ms / prof_steps),
                   "wall_ms_per_step_incl_flush": wall / args.steps * 1e3},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": 208, "d2h_bytes_per_step": int(npix * 16 + 208 * (n_iter / args.steps / world)),
                "note": "api.Scene.render() into pinned host memory: film clear, waves, resolve, D2H of the W*H float4 image; the scene (BVH, tables) stays resident like model weights; H2D per step is the wave control block"},
        "gpu_launches": int(n_launch + args.steps * (3 + 1)),
        "events": {"n_gen": n_gen, "n_closest": n_closest, "n_hit": n_hit, "n_next": n_next, "n_shadow": n_shadow, "n_shadow_miss": n_shadow_miss,
                   "n_splat": n_splat, "iterations_per_step": n_iter / args.steps / world},
        "roofline": roofline,
        "roofline_stages": stages,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter on stdout (e.g. the "NCCL version ..." banner) must not pollute the one JSON line: fd 1 is pointed at
    stderr for the whole run and the result is written to the original stdout at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--state-count", type=int, default=8388608,
                    help="path-state pool (the renderer's `state_count` parameter); 8 Mi states = 2.4 GB of queues (profiles/r1h_sweep_pool.log)")
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"],
                    help="multi-GPU split: sample-index ranges (default, load balanced) or image tiles (row bands, BASELINE config 5's wording)")
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel (profiling runs); reported in config")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (roofline bvh statistics, cpu_baseline): profiler runs")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
