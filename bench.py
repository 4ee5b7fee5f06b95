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
DEFAULT_WORKLOAD = "c2_disney_1080p_256spp"   # --gpus 1: the configuration the metric is quoted on
SCALING_WORKLOAD = "c5_disney_4k_1024spp"     # --gpus N > 1: BASELINE configs[4], total work fixed (strong scaling)


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


def _partition_text(tiles, world, W, H):
    from btrc_b200 import partition as P
    if not tiles:
        return "sample-index ranges per rank"
    rows = P.strip_rows(world, H)
    shape = "%d x %d pixel tiles" % (P.tile_width(world, W), rows) if P.tiles_are_2d(world, W, H, rows) else "strips of %d pixel rows" % rows
    return "image tiles: " + shape + " dealt round-robin to the ranks (tile-mode indexing)"


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
BVH_ALG_CACHE = os.path.join(ROOT, "profiles", "bvh_alg.json")


def bvh_alg_per_ray(scene_path, sample_paths=16384, workload=None):
    """lib/bvh binary-BVH statistics: mean traversal steps and primitive tests per closest-hit ray, measured
    by oracle/_ref (the reference's vendored SingleRayTraverser, sweep-SAH) on the ray batches of a bounded path sample.
    The workload scenes are deterministic, so are these statistics: profiles/bvh_alg.json holds them per workload (written by
    `python bench.py --write-bvh-alg`, the same function on the same scenes) and saves the default run the 10-million-triangle
    CPU build; a workload that is not in the file is measured on the spot."""
    import ctypes

    if workload and os.path.exists(BVH_ALG_CACHE):
        hit = json.load(open(BVH_ALG_CACHE)).get(workload)
        if hit:
            return dict(hit, source="profiles/bvh_alg.json")

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


_WRITE_ONLY = []


def write_only_gbs():
    """GB/s of a pure-write stream on this GPU (torch fill of 2 GiB, CUDA events): the ceiling of a kernel that only writes."""
    if not _WRITE_ONLY:
        import torch
        n = 2 << 30
        buf = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            buf.zero_()
        e1.record()
        torch.cuda.synchronize()
        _WRITE_ONLY.append(5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del buf
    return _WRITE_ONLY[0]


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
class Dist:
    """Process-group facts + the renderer's own NCCL communicator (created once, shared by every workload of the run)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        from btrc_b200 import api

        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        api.set_device(self.local_rank)
        self.comm = None
        if self.world > 1:
            # torch.distributed is the plumbing (rendezvous, barrier, max over ranks); the data path's one collective -- the film
            # reduce -- runs inside the library on its own communicator (btrc_reduce_films), whose id travels over this group
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            box = [api.nccl_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.comm = api.Comm(self.world, self.rank, box[0])

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        import torch
        import torch.distributed as dist

        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, xs):
        import torch
        import torch.distributed as dist

        t = torch.tensor(list(xs), device="cuda", dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def close(self):
        import torch.distributed as dist

        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            dist.destroy_process_group()


def dram_traffic(workload):
    """DRAM bytes per ray / state of the traversal and shade kernels from the committed `ncu --set full` capture of this
    workload (profiles/traffic_<workload>.json, written by scripts/ncu_summary.py traffic)."""
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)
    return json.load(open(tpath)) if os.path.exists(tpath) else {}


def measure(args, D, workload, steps, warmup, spp_override=0, e2e=True, cpu=True, scaling=None, partition=None, bvh_stats=True):
    """Times `steps` renders of `workload` on the process group D and returns rank 0's result line (None on other ranks)."""
    import torch

    from btrc_b200 import api

    rank, world = D.rank, D.world
    w = WORKLOADS[workload]
    W, H = w["width"], w["height"]
    scaling = scaling or ("strong" if w.get("strong") else "weak")
    partition = partition or ("tiles" if w.get("strong") and world > 1 and args.partition == "auto" else (args.partition if args.partition != "auto" else "samples"))
    tiles = partition == "tiles"
    spp_total = spp_override or w["spp"]
    if scaling == "weak":        # per-GPU work fixed: the scene holds spp * world samples per pixel, every rank renders `spp` of them
        spp_scene = spp_total * world
    else:                        # total work fixed
        spp_scene = spp_total
        if not tiles and spp_scene % world:
            raise SystemExit(f"bench.py: {spp_scene} spp do not split evenly over {world} ranks")
    npix = W * H
    tmp = tempfile.TemporaryDirectory()
    path = make_scene(workload, tmp.name, spp=spp_scene, tile=tiles)
    s = api.Scene(path, state_count=args.state_count)
    s.set_profiling(False)  # the timed steps carry no per-stage events; the stage breakdown comes from a separate profiled pass
    stream = torch.cuda.ExternalStream(s.stream_handle())
    color_dev = torch.zeros(H, W, 4, device="cuda") if rank == 0 else None
    color_host = torch.zeros(H, W, 4).pin_memory() if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        """One device-resident step: film clear -> this rank's share of the waves -> [ncclReduce of the films] -> resolve on rank 0."""
        s.film_clear()
        st = s.render_part(world, rank)
        if world > 1:
            s.reduce_films(D.comm, 0)
        if rank == 0:
            s.resolve(color_dev)
        return st

    def step_e2e():
        """The same step through the public call with HOST buffers: result image read back to pinned host memory."""
        if world == 1:
            _, _, _, st = s.render(color_host.numpy())
            return st
        s.film_clear()
        st = s.render_part(world, rank)
        s.reduce_films(D.comm, 0)
        if rank == 0:
            s.resolve(color_host.numpy())
        return st

    # NVML is initialised before the warm-up steps (its start-up takes driver locks that stall launches for ~100 ms)
    sampler = ClockSampler(D.local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(warmup, 3)):
        step()
    if rank == 0:
        sampler.reset()

    # ---- timed region: K device-resident steps, CUDA events on the renderer's stream, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    agg = None
    D.barrier()
    t_wall0 = time.perf_counter()
    for k in range(steps):
        flush.fill_(k & 0xff)
        torch.cuda.synchronize()
        ev[k][0].record(stream)
        st = step()
        ev[k][1].record(stream)
        if agg is None:
            agg = {key: 0 for key in st if key != "stage_ms"}
        for key, v in st.items():
            if key != "stage_ms":
                agg[key] += v
    D.barrier()
    wall = time.perf_counter() - t_wall0
    dev_ms = D.max(sum(a.elapsed_time(b) for a, b in ev))
    render_ms_max = D.max(agg["render_ms"])           # slowest rank's time inside the waves: what the partition balances

    # ---- stage breakdown: the same steps again with one CUDA event between consecutive kernels (not part of `value`)
    s.set_profiling(True)
    prof_steps = min(steps, 3)
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
    e2e_s = None
    if e2e:
        D.barrier()
        e2e_t0 = time.perf_counter()
        for k in range(steps):
            step_e2e()
        D.barrier()
        e2e_s = D.max(time.perf_counter() - e2e_t0)
    clocks = sampler.stop() if rank == 0 else None

    # whole-job event counts
    keys = ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat", "n_iter", "n_launches")
    n_gen, n_closest, n_hit, n_next, n_shadow, n_shadow_miss, n_splat, n_iter, n_launch = D.sum(agg[k] for k in keys)
    render_ms_sum = D.sum([agg["render_ms"]])[0]
    if rank != 0:
        s.close()
        return None
    secs = dev_ms / 1e3
    value = n_gen / secs / 1e6
    mrays = (n_closest + n_shadow) / secs / 1e6

    # ---- roofline of the dominant kernel (rank 0's stage times; algorithmic bytes per SURVEY 8d)
    bvh = bvh_alg_per_ray(path, workload=workload) if bvh_stats else None   # per ray: independent of the spp override
    peak, peak_src = measured_peak()
    sb = stage_bytes({k: agg[k] for k in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat")}, bvh,
                     has_medium=bool(s.info.has_medium))
    iters = max(prof_iters, 1)
    stages = {}
    sb["bookkeeping"] = 0
    scale = prof_steps / steps  # the byte totals cover K timed steps, the stage times `prof_steps` identical ones
    traffic_db = dram_traffic(workload)
    units_of = {"trace": agg["n_closest"], "shadow": agg["n_shadow"], "shade": agg["n_hit"], "medium": agg["n_closest"], "generate": agg["n_gen"]}
    for name in ("generate", "trace", "medium", "shade", "shadow", "bookkeeping"):
        ms = stage_ms[name]
        if name == "medium" and not s.info.has_medium:
            continue
        nbytes = sb[name] * scale
        gbs = nbytes / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        st_line = {"ms_total": ms, "launches": iters, "ms_per_launch": ms / iters, "alg_bytes_per_launch": nbytes / iters,
                   "achieved_gbs": gbs, "frac": gbs / peak, "share_of_step": ms / prof_ms}
        t = traffic_db.get(name)
        if t and t.get("dram_bytes_per_unit") and ms > 0:
            dram = t["dram_bytes_per_unit"] * units_of[name] * scale
            # measured DRAM bytes (ncu, per ray / state) x this run's units over this run's stage time: the fraction of the HBM
            # peak the kernel really moves, next to the SURVEY 8d figure whose BVH term is served by L1 / L2 when the tree fits
            st_line["dram_gbs"] = dram / (ms / 1e3) / 1e9
            st_line["dram_frac"] = st_line["dram_gbs"] / peak
            st_line["dram_bytes_per_launch"] = dram / iters
        stages[name] = st_line
    if "generate" in stages and stages["generate"]["ms_total"] > 0:
        # generate writes 104 B per path and reads nothing.  A write-only stream does not reach the copy figure of MEASURED_PEAKS.json
        # (read + write bytes): measure what it does reach here (a library fill of 2 GiB, outside every timed region) as the yardstick.
        wo = write_only_gbs()
        stages["generate"]["write_only_peak_gbs"] = wo
        stages["generate"]["frac_of_write_only_peak"] = stages["generate"]["achieved_gbs"] / wo
    dom = max((k for k in stages if k != "bookkeeping"), key=lambda k: stages[k]["ms_total"])
    info = s.info
    bvh_bytes = int(info.bvh_bytes) if hasattr(info, "bvh_bytes") else None
    L2 = 126 << 20
    resident = bvh_bytes is not None and bvh_bytes < L2
    roofline = {"kernel": dom + "_kernel",
                "bound": "hbm", "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                "traffic": stages[dom].get("dram_bytes_per_launch"),
                "traffic_source": ("profiles/traffic_%s.json (ncu dram__bytes_read+write per ray x rays per launch)" % workload) if stages[dom].get("dram_bytes_per_launch") else None,
                "dram_frac": stages[dom].get("dram_frac"),
                "limiter": ("L1/L2 latency and SIMT divergence: the BVH (%.0f MB) is resident in the 126 MB L2, so `frac` (SURVEY 8d algorithmic bytes, 96 %% of them "
                            "BVH nodes) is the rate at which the caches serve the walk, `dram_frac` what reaches HBM" % (bvh_bytes / 1e6)) if resident else
                           ("HBM latency: the BVH (%.0f MB) exceeds the L2" % (bvh_bytes / 1e6) if bvh_bytes else None),
                "peak_source": peak_src, "alg_bytes_per_launch": stages[dom]["alg_bytes_per_launch"],
                "ms_per_launch": stages[dom]["ms_per_launch"], "bvh_alg": bvh}

    # ---- CPU baseline: the reference's own CPU execution of the path (CpuArm) on the host cores of this box, bounded sample
    cpu_line = None
    if cpu:
        arm = CpuArm(make_scene(workload, os.path.join(tmp.name, "cpu")))
        t1, _ = arm.render(1)  # calibration pass: 1 spp
        cpu_spp = int(max(1, min(w["spp"], 15.0 / max(t1, 1e-3))))  # ~15 s of CPU work
        cpu_s, _ = arm.render(cpu_spp)
        cpu_line = {"value": npix * cpu_spp / cpu_s / 1e6, "unit": "Msamples/s", "cores": arm.cores, "kind": arm.kind,
                    "sample": f"{cpu_spp} of {w['spp']} spp of the workload ({W}x{H}), {cpu_s:.1f} s on all host cores; " + arm.note}

    cfg = workload_config(workload)   # identical in both arms; how this arm ran it goes to "run"
    run = {}
    run.update({"spp_rendered_per_step": spp_scene, "gpus": world, "state_count": s.info.state_count, "bvh_layout": "instanced" if s.info.bvh_instanced else "wide8" if s.info.bvh_wide else "flat",
                "partition": (_partition_text(tiles, world, W, H) +
                              "; ncclReduce(sum) of the radiance+weight films onto rank 0 inside the library (btrc_reduce_films), resolve on rank 0") if world > 1 else "single GPU",
                "timing": "CUDA events on the renderer's stream per step, max over ranks; 256 MB L2 flush (untimed) between steps; "
                          "roofline_stages from a separate pass of %d identical steps with one CUDA event between kernels (%.1f ms/step)" % (prof_steps, prof_ms / prof_steps),
                "wall_ms_per_step_incl_flush": wall / steps * 1e3})
    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "mrays_per_s": mrays, "config": cfg, "run": run, "clocks": clocks,
        "gpu_launches": int(n_launch + steps * (3 + 1)),
        "events": {"n_gen": n_gen, "n_closest": n_closest, "n_hit": n_hit, "n_next": n_next, "n_shadow": n_shadow, "n_shadow_miss": n_shadow_miss,
                   "n_splat": n_splat, "iterations_per_step": n_iter / steps / world},
        "roofline": roofline,
        "roofline_stages": stages,
        "cpu_baseline": cpu_line,
    }
    if e2e_s is not None:
        ctl_bytes = api.wave_ctl_bytes()
        line["e2e"] = {"value": n_gen / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": ctl_bytes,
                       "d2h_bytes_per_step": int(npix * 16 + ctl_bytes * (n_iter / steps / world)),
                       "note": "api.Scene.render() into pinned host memory: film clear, waves, resolve, D2H of the W*H float4 image; the scene (BVH, tables) stays "
                               "resident like model weights; H2D per step is the wave control block"}
    if world > 1:
        # what limits the curve: the slowest rank's wave time against the mean (partition balance), and what is left of the step
        # after the waves (film reduce + resolve + launch gaps)
        line["multi_gpu"] = {"render_ms_per_step_slowest_rank": render_ms_max / steps, "render_ms_per_step_mean_rank": render_ms_sum / steps / world,
                             "imbalance": render_ms_max / (render_ms_sum / world) - 1.0,
                             "reduce_resolve_and_gaps_ms_per_step": dev_ms / steps - render_ms_max / steps}
    s.close()
    return line


def run_ours(args):
    D = Dist()
    workload = args.workload or (DEFAULT_WORKLOAD if D.world == 1 else SCALING_WORKLOAD)
    line = measure(args, D, workload, args.steps, args.warmup, spp_override=args.spp, cpu=not args.no_cpu_baseline, bvh_stats=not args.no_cpu_baseline)
    if D.world == 1 and not args.spp and not args.no_workloads and args.workload is None:
        # the other BASELINE configurations, short runs (value, clocks, per-stage shares and roofline fractions): BENCH carries all five
        extra = {}
        for wl, k, spp in (("c1_cornell_512_16spp", 200, 0), ("c3_medium_1080p_128spp", 3, 0), ("c4_instanced_10m_1080p_128spp", 2, 0),
                           ("c5_disney_4k_1024spp", 2, 128)):
            try:
                # the medium workload peaks at a 16 Mi pool (profiles/r2j_sweep_pool.md: 747 vs 716 Msamples/s at 32 Mi)
                pool = args.state_count
                if wl == "c3_medium_1080p_128spp":
                    args.state_count = min(pool, 16777216)
                try:
                    r = measure(args, D, wl, k, 3, spp_override=spp, e2e=False, cpu=False, scaling="weak")
                finally:
                    args.state_count = pool
                extra[wl] = {"value": r["value"], "unit": "Msamples/s", "mrays_per_s": r["mrays_per_s"], "ms_per_step": r["ms_per_step"], "steps": k,
                             "spp_rendered_per_step": r["run"]["spp_rendered_per_step"], "clocks": r["clocks"],
                             "stages": {n: {"share_of_step": v["share_of_step"], "frac": v["frac"], "dram_frac": v.get("dram_frac"), "ms_per_launch": v["ms_per_launch"]}
                                        for n, v in r["roofline_stages"].items()}}
                if wl == "c5_disney_4k_1024spp":
                    extra[wl]["note"] = "one GPU's share (128 spp) of the 8-GPU strong-scaling run of this configuration: `bench.py --gpus 8` renders all 1024 spp"
            except Exception as e:  # a side measurement must not take the headline line with it
                extra[wl] = {"error": str(e)[:200]}
        if line is not None:
            line["workloads"] = extra
    if line is not None:
        emit(line)
    D.close()


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
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: %s on one GPU, %s (strong scaling) on several" % (DEFAULT_WORKLOAD, SCALING_WORKLOAD))
    ap.add_argument("--write-bvh-alg", action="store_true", help="(CPU only) measure the lib/bvh statistics of every workload scene into profiles/bvh_alg.json")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short side runs of the other BASELINE configurations")
    ap.add_argument("--state-count", type=int, default=33554432,
                    help="path-state pool (the renderer's `state_count` parameter); 32 Mi states = 9.4 GB of queues.  C2 with the block-coherent "
                         "generation order: 1192 / 1242 / 1273 Msamples/s at 8 / 16 / 32 Mi (profiles/r2j_sweep_pool.md)")
    ap.add_argument("--partition", default="auto", choices=["auto", "samples", "tiles"],
                    help="multi-GPU split: sample-index ranges or image tiles (64-pixel-wide tiles or strips of pixel rows, dealt round-robin); auto = tiles for the "
                         "strong-scaling configuration (BASELINE config 5's wording), samples otherwise")
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel (profiling runs); reported in config")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (roofline bvh statistics, cpu_baseline): profiler runs")
    args = ap.parse_args()
    if args.write_bvh_alg:
        out = {}
        for wl in sorted(WORKLOADS):
            with tempfile.TemporaryDirectory() as d:
                out[wl] = bvh_alg_per_ray(make_scene(wl, d, spp=1))
            print(wl, out[wl], file=sys.stderr)
        json.dump(out, open(BVH_ALG_CACHE, "w"), indent=1)
        return
    quiet_stdout()
    if args.impl == "reference":
        if args.workload is None:
            args.workload = DEFAULT_WORKLOAD if int(os.environ.get("WORLD_SIZE", "1")) == 1 else SCALING_WORKLOAD
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
