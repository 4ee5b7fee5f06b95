"""ctypes binding of libbtrc_b200.so (include/btrc_b200.h) -- the call a user of the render path makes.

Mirrors the three calls of the reference CLI (src/cli/main.cpp:36-80): load (parse + create + commit),
render, run post-processors; plus the stage-level entry points used by the parity tests.  Arrays are numpy
(host) or anything exposing a CUDA pointer via `.data_ptr()` (torch tensors).  There is no CPU fallback: if the
library is missing this module raises, and device calls fail with the library's CUDA error when no GPU exists.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BTRC_B200_LIB") or os.path.join(_HERE, "libbtrc_b200.so")  # the override serves A/B builds of the same library

HOST, DEVICE = 0, 1


class BtrcError(RuntimeError):
    """Chained error text from the library (one context per line, like the reference's nested BtrcException)."""


class SceneInfo(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("width", "height", "spp", "min_depth", "max_depth", "state_count", "tile", "albedo", "normal",
                                       "n_instances", "n_lights", "n_triangles", "n_materials", "bvh_depth", "n_media", "has_medium", "n_textures")] + \
               [("world_diagonal", c_float), ("bvh_build_seconds", c_float), ("bvh_bytes", c_uint64), ("bvh_instanced", c_int32), ("bvh_wide", c_int32)]


class RenderStats(ctypes.Structure):
    _fields_ = [(n, c_uint64) for n in ("n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat", "n_iter")] + \
               [("render_ms", c_double), ("stage_ms", c_double * 6), ("n_launches", c_uint64)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["stage_ms"] = dict(zip(("generate", "trace", "shade", "shadow", "medium", "bookkeeping"), [float(x) for x in self.stage_ms]))
        return d


_lib = None
PROGRESS_FN = ctypes.CFUNCTYPE(c_int, ctypes.c_float, c_void_p)
BTRC_STOPPED = 2


class RenderStopped(RuntimeError):
    """The reporter asked the renderer to stop (Renderer::should_stop): no image was produced."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BtrcError(f"{LIB_PATH} is missing: run `python -m btrc_b200.build` (the render path has no fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.btrc_last_error.restype = c_char_p
        L.btrc_version.restype = c_char_p
        L.btrc_scene_load_json.argtypes = [c_char_p, POINTER(c_void_p)]
        L.btrc_scene_load_json_string.argtypes = [c_char_p, c_char_p, POINTER(c_void_p)]
        L.btrc_scene_free.argtypes = [c_void_p]
        L.btrc_scene_free.restype = None
        L.btrc_scene_info_get.argtypes = [c_void_p, POINTER(SceneInfo)]
        L.btrc_set_params.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32]
        L.btrc_render.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, POINTER(RenderStats)]
        L.btrc_film_clear.argtypes = [c_void_p]
        L.btrc_stream.argtypes = [c_void_p]
        L.btrc_stream.restype = c_void_p
        L.btrc_set_profiling.argtypes = [c_void_p, c_int32]
        L.btrc_render_range.argtypes = [c_void_p, c_int64, c_int64, POINTER(RenderStats)]
        L.btrc_film_export.argtypes = [c_void_p, c_int] + [c_void_p] * 4
        L.btrc_film_import.argtypes = [c_void_p, c_int] + [c_void_p] * 4
        L.btrc_resolve.argtypes = [c_void_p, c_int] + [c_void_p] * 3
        L.btrc_run_post_processors.argtypes = [c_void_p, c_int] + [c_void_p] * 3
        L.btrc_closest_hit.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
        L.btrc_any_hit.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int32, c_void_p]
        L.btrc_generate.argtypes = [c_void_p, c_int, c_int64, c_int32] + [c_void_p] * 7
        L.btrc_shade.argtypes = [c_void_p, c_int, c_int32] + [c_void_p] * 9 + [POINTER(c_int32)] + [c_void_p] * 7 + [POINTER(c_int32)] + [c_void_p] * 4
        L.btrc_medium.argtypes = L.btrc_shade.argtypes
        L.btrc_shadow.argtypes = [c_void_p, c_int, c_int32] + [c_void_p] * 4
        L.btrc_geometry_tables.argtypes = [c_void_p, c_int32, POINTER(c_int32), c_void_p]
        L.btrc_geometry_positions.argtypes = [c_void_p, c_int32, POINTER(c_int32), c_void_p]
        L.btrc_instance_transform.argtypes = [c_void_p, c_int32, c_void_p, c_void_p]
        L.btrc_instance_geometry.argtypes = [c_void_p, c_int32, POINTER(c_int32)]
        L.btrc_ibl_table.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32), c_void_p]
        L.btrc_set_reporter.argtypes = [c_void_p, PROGRESS_FN, c_void_p]
        L.btrc_image_load.argtypes = [c_char_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), c_void_p, c_uint64]
        L.btrc_scene_file_dump.argtypes = [c_char_p, c_void_p, c_uint64, POINTER(c_uint64)]
        L.btrc_scene_file_parse.argtypes = [c_char_p, c_char_p, c_int32, c_void_p, POINTER(c_int32)]
        L.btrc_device_sm_count.restype = c_int
        L.btrc_device_count.restype = c_int
        L.btrc_wave_ctl_bytes.restype = c_int
        L.btrc_set_device.argtypes = [c_int32]
        L.btrc_render_part.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(RenderStats)]
        L.btrc_nccl_unique_id.argtypes = [c_void_p]
        L.btrc_comm_create.argtypes = [c_int32, c_int32, c_void_p, POINTER(c_void_p)]
        L.btrc_comm_destroy.argtypes = [c_void_p]
        L.btrc_comm_destroy.restype = None
        L.btrc_reduce_films.argtypes = [c_void_p, c_void_p, c_int32]
        _lib = L
    return _lib


def _check(rc):
    if rc == BTRC_STOPPED:
        raise RenderStopped(lib().btrc_last_error().decode())
    if rc != 0:
        raise BtrcError(lib().btrc_last_error().decode())


def _p(a):
    """Pointer of a numpy array (host) or a torch tensor (device or host); None stays None."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"]
    return c_void_p(a.ctypes.data)


def _img(a, what, count, dtype="float32"):
    """Caller-supplied image / film buffers go to the library as raw pointers: refuse anything whose element type or size is
    not what the entry point will read or write (numpy arrays and torch tensors)."""
    if a is None:
        return None
    n = a.numel() if hasattr(a, "numel") else a.size
    dt = str(a.dtype).replace("torch.", "")
    if dt != dtype:
        raise BtrcError("%s: expected %s elements, got %s" % (what, dtype, dt))
    if n != count:
        raise BtrcError("%s: expected %d elements, got %d" % (what, count, n))
    if hasattr(a, "is_contiguous") and not a.is_contiguous():
        raise BtrcError("%s: tensor is not contiguous" % what)
    return a


def _mem(*arrays):
    for a in arrays:
        if a is not None and hasattr(a, "is_cuda") and a.is_cuda:
            return DEVICE
    return HOST


def new_state(n):
    """Path-state batch in the reference SoA layout (wavefront/soa.h:8-37)."""
    f32 = np.float32
    return dict(o_med=np.zeros((n, 4), f32), d_t1=np.zeros((n, 4), f32), pixel=np.zeros((n, 2), np.uint32),
                beta_depth=np.zeros((n, 4), f32), path_rad=np.zeros((n, 4), f32), rng=np.zeros((n, 2), np.uint64),
                ble=np.zeros((n, 4), f32))


_STATE_KEYS = ("o_med", "d_t1", "pixel", "beta_depth", "path_rad", "rng", "ble")
_STATE_TYPES = dict(o_med=(np.float32, 4), d_t1=(np.float32, 4), pixel=(np.uint32, 2), beta_depth=(np.float32, 4), path_rad=(np.float32, 4),
                    rng=(np.uint64, 2), ble=(np.float32, 4))


def _state(st, copy=False):
    """A caller's path-state batch with the element types and row widths the C ABI reads (raw pointers carry neither)."""
    n = np.asarray(st["o_med"]).shape[0]
    out = {}
    for k in _STATE_KEYS:
        ty, w = _STATE_TYPES[k]
        a = np.asarray(st[k])
        if a.dtype != ty or a.shape != (n, w):
            raise BtrcError("state array %r: expected %s[%d, %d], got %s%r" % (k, np.dtype(ty).name, n, w, a.dtype, a.shape))
        out[k] = np.array(a, order="C", copy=True) if copy else np.ascontiguousarray(a)
    return out


class Scene:
    """A loaded + committed scene and its wavefront renderer (reference: Scene + WavefrontPathTracer)."""

    def __init__(self, path=None, json_text=None, scene_dir=None, spp=None, state_count=None, min_depth=None, max_depth=None):
        self.L = lib()
        self.h = c_void_p()
        if path is not None:
            _check(self.L.btrc_scene_load_json(os.fsencode(path), byref(self.h)))
        else:
            _check(self.L.btrc_scene_load_json_string(json_text.encode(), os.fsencode(scene_dir or "."), byref(self.h)))
        if any(v is not None for v in (spp, state_count, min_depth, max_depth)):
            self.set_params(spp, state_count, min_depth, max_depth)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.btrc_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, spp=None, state_count=None, min_depth=None, max_depth=None):
        v = [(-1 if x is None else int(x)) for x in (spp, state_count, min_depth, max_depth)]
        _check(self.L.btrc_set_params(self.h, *v))

    @property
    def info(self) -> SceneInfo:
        i = SceneInfo()
        _check(self.L.btrc_scene_info_get(self.h, byref(i)))
        return i

    # ---- whole render -------------------------------------------------------------------------------------
    def render(self, color=None, albedo=None, normal=None):
        """Renderer::render(): returns (color, albedo, normal, stats); arrays are (H, W, 4) float32."""
        i = self.info
        if color is None:
            color = np.zeros((i.height, i.width, 4), np.float32)
        if albedo is None and i.albedo:
            albedo = np.zeros((i.height, i.width, 4), np.float32)
        if normal is None and i.normal:
            normal = np.zeros((i.height, i.width, 4), np.float32)
        n4 = 4 * i.width * i.height
        _img(color, "render color", n4); _img(albedo, "render albedo", n4); _img(normal, "render normal", n4)
        st = RenderStats()
        _check(self.L.btrc_render(self.h, _mem(color), _p(color), _p(albedo), _p(normal), byref(st)))
        return color, albedo, normal, st.as_dict()

    def set_reporter(self, fn=None):
        """Reporter::progress + Renderer::should_stop: fn(percentage) is called from the rendering thread; a truthy return
        value stops the render (render() / render_range() then raise RenderStopped).  None removes the reporter."""
        if fn is None:
            self._reporter = None
            _check(self.L.btrc_set_reporter(self.h, ctypes.cast(None, PROGRESS_FN), None))
            return
        self._reporter = PROGRESS_FN(lambda pct, _user: 1 if fn(pct) else 0)  # kept alive with the scene
        _check(self.L.btrc_set_reporter(self.h, self._reporter, None))

    def set_profiling(self, on=True):
        _check(self.L.btrc_set_profiling(self.h, 1 if on else 0))

    def stream_handle(self) -> int:
        """cudaStream_t of the renderer (wrap with torch.cuda.ExternalStream to record events on it)."""
        p = self.L.btrc_stream(self.h)
        if not p:
            raise BtrcError(self.L.btrc_last_error().decode())
        return int(p)

    def film_clear(self):
        _check(self.L.btrc_film_clear(self.h))

    def render_range(self, begin, end):
        st = RenderStats()
        _check(self.L.btrc_render_range(self.h, int(begin), int(end), byref(st)))
        return st.as_dict()

    def _film_args(self, radiance, weight, albedo, normal):
        i = self.info
        npix = i.width * i.height
        _img(radiance, "film radiance", 4 * npix); _img(weight, "film weight", npix)
        _img(albedo, "film albedo", 4 * npix); _img(normal, "film normal", 4 * npix)

    def render_part(self, n_parts, part, rows_per_strip=0):
        """This device's share of the render (sample ranges, or interleaved image tiles when the renderer is in tile mode: partition.tile_owner)."""
        st = RenderStats()
        _check(self.L.btrc_render_part(self.h, int(n_parts), int(part), int(rows_per_strip), byref(st)))
        return st.as_dict()

    def reduce_films(self, comm, root=0):
        """ncclReduce(sum) of the film planes onto `root`, on the renderer's stream (comm: Comm)."""
        _check(self.L.btrc_reduce_films(self.h, comm.h, int(root)))

    def film_export(self, radiance, weight, albedo=None, normal=None):
        self._film_args(radiance, weight, albedo, normal)
        _check(self.L.btrc_film_export(self.h, _mem(radiance, weight), _p(radiance), _p(weight), _p(albedo), _p(normal)))

    def film_import(self, radiance, weight, albedo=None, normal=None):
        self._film_args(radiance, weight, albedo, normal)
        _check(self.L.btrc_film_import(self.h, _mem(radiance, weight), _p(radiance), _p(weight), _p(albedo), _p(normal)))

    def resolve(self, color, albedo=None, normal=None):
        n4 = 4 * self.info.width * self.info.height
        _img(color, "resolve color", n4); _img(albedo, "resolve albedo", n4); _img(normal, "resolve normal", n4)
        _check(self.L.btrc_resolve(self.h, _mem(color), _p(color), _p(albedo), _p(normal)))

    def run_post_processors(self, color, albedo=None, normal=None):
        _check(self.L.btrc_run_post_processors(self.h, _mem(color), _p(color), _p(albedo), _p(normal)))

    # ---- stage level ----------------------------------------------------------------------------------------
    def closest_hit(self, o_med, d_t1):
        n = o_med.shape[0]
        flag = np.zeros(n, np.uint32)
        detail = np.zeros((n, 4), np.uint32)
        _check(self.L.btrc_closest_hit(self.h, HOST, _p(np.ascontiguousarray(o_med, np.float32)), _p(np.ascontiguousarray(d_t1, np.float32)), n, _p(flag), _p(detail)))
        return flag, detail

    def any_hit(self, o_med, d_t1):
        n = o_med.shape[0]
        occ = np.zeros(n, np.uint8)
        _check(self.L.btrc_any_hit(self.h, HOST, _p(np.ascontiguousarray(o_med, np.float32)), _p(np.ascontiguousarray(d_t1, np.float32)), n, _p(occ)))
        return occ

    def generate(self, first, n):
        st = new_state(n)
        _check(self.L.btrc_generate(self.h, HOST, int(first), n, *[_p(st[k]) for k in _STATE_KEYS]))
        return st

    def shade(self, st, flag, detail):
        st = _state(st)
        n = st["o_med"].shape[0]
        nxt = new_state(n)
        sh = dict(pixel=np.zeros((n, 2), np.uint32), beta_li=np.zeros((n, 4), np.float32), o_med=np.zeros((n, 4), np.float32),
                  d_t1=np.zeros((n, 4), np.float32))
        n_next, n_shadow = c_int32(0), c_int32(0)
        flag = np.ascontiguousarray(flag, np.uint32)
        detail = np.ascontiguousarray(detail, np.uint32)
        _check(self.L.btrc_shade(self.h, HOST, n, *[_p(st[k]) for k in _STATE_KEYS], _p(flag), _p(detail), byref(n_next),
                                 *[_p(nxt[k]) for k in _STATE_KEYS], byref(n_shadow), _p(sh["pixel"]), _p(sh["beta_li"]), _p(sh["o_med"]), _p(sh["d_t1"])))
        nn, ns = n_next.value, n_shadow.value
        return {k: v[:nn] for k, v in nxt.items()}, {k: v[:ns] for k, v in sh.items()}

    def medium(self, st, flag, detail):
        """Medium stage; returns (state updated in place, flag with the scattered bit, next states, shadow rays)."""
        st = _state(st, copy=True)
        n = st["o_med"].shape[0]
        flag = np.ascontiguousarray(flag.copy(), np.uint32)
        detail = np.ascontiguousarray(detail, np.uint32)
        nxt = new_state(n)
        sh = dict(pixel=np.zeros((n, 2), np.uint32), beta_li=np.zeros((n, 4), np.float32), o_med=np.zeros((n, 4), np.float32),
                  d_t1=np.zeros((n, 4), np.float32))
        n_next, n_shadow = c_int32(0), c_int32(0)
        _check(self.L.btrc_medium(self.h, HOST, n, *[_p(st[k]) for k in _STATE_KEYS], _p(flag), _p(detail), byref(n_next),
                                  *[_p(nxt[k]) for k in _STATE_KEYS], byref(n_shadow), _p(sh["pixel"]), _p(sh["beta_li"]), _p(sh["o_med"]), _p(sh["d_t1"])))
        nn, ns = n_next.value, n_shadow.value
        return st, flag, {k: v[:nn] for k, v in nxt.items()}, {k: v[:ns] for k, v in sh.items()}

    def shadow(self, sh):
        n = sh["pixel"].shape[0]
        _check(self.L.btrc_shadow(self.h, HOST, n, _p(np.ascontiguousarray(sh["pixel"], np.uint32)), _p(np.ascontiguousarray(sh["beta_li"], np.float32)),
                                  _p(np.ascontiguousarray(sh["o_med"], np.float32)), _p(np.ascontiguousarray(sh["d_t1"], np.float32))))

    def geometry_tables(self, geometry_id):
        n = c_int32(0)
        _check(self.L.btrc_geometry_tables(self.h, geometry_id, byref(n), None))
        out = np.zeros((n.value, 24), np.float32)
        _check(self.L.btrc_geometry_tables(self.h, geometry_id, byref(n), _p(out)))
        return out

    def geometry_positions(self, geometry_id):
        """Local-space corner positions, (ntri, 3, 3) float32 (host side only)."""
        n = c_int32(0)
        _check(self.L.btrc_geometry_positions(self.h, geometry_id, byref(n), None))
        out = np.zeros((n.value, 3, 3), np.float32)
        _check(self.L.btrc_geometry_positions(self.h, geometry_id, byref(n), _p(out)))
        return out

    def ibl_table(self):
        w, h = c_int32(0), c_int32(0)
        _check(self.L.btrc_ibl_table(self.h, byref(w), byref(h), None))
        t = np.zeros((h.value, w.value), np.float32)
        _check(self.L.btrc_ibl_table(self.h, byref(w), byref(h), _p(t)))
        return t

    def instance_geometry(self, instance_id):
        g = c_int32(0)
        _check(self.L.btrc_instance_geometry(self.h, int(instance_id), byref(g)))
        return g.value

    def instance_transform(self, instance_id):
        mat = np.zeros((3, 4), np.float32)
        inv = np.zeros((3, 4), np.float32)
        _check(self.L.btrc_instance_transform(self.h, instance_id, _p(mat), _p(inv)))
        return mat, inv


def nccl_unique_id() -> bytes:
    """128 bytes of an ncclUniqueId (create on one rank, hand to every rank's Comm)."""
    buf = ctypes.create_string_buffer(128)
    _check(lib().btrc_nccl_unique_id(buf))
    return buf.raw


class Comm:
    """The renderer's own NCCL communicator (one rank per process / GPU)."""

    def __init__(self, n_ranks, rank, unique_id: bytes | None):
        self.h = c_void_p()
        buf = ctypes.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check(lib().btrc_comm_create(int(n_ranks), int(rank), buf, byref(self.h)))

    def close(self):
        if self.h:
            lib().btrc_comm_destroy(self.h)
            self.h = c_void_p()


def wave_ctl_bytes():
    """Bytes of the device-resident wave control block (uploaded once per render, read back per progress snapshot)."""
    return int(lib().btrc_wave_ctl_bytes())


def set_device(index):
    _check(lib().btrc_set_device(int(index)))


def load_image(path):
    """ImageDynamic::load: (H, W, C) uint8 for .png / .jpg, (H, W, 3) float32 for .hdr."""
    L = lib()
    w, h, c, isf = c_int32(0), c_int32(0), c_int32(0), c_int32(0)
    _check(L.btrc_image_load(path.encode(), byref(w), byref(h), byref(c), byref(isf), None, 0))
    out = np.zeros((h.value, w.value, c.value), np.float32 if isf.value else np.uint8)
    _check(L.btrc_image_load(path.encode(), byref(w), byref(h), byref(c), byref(isf), out.ctypes.data, out.nbytes))
    return out


def scene_file_dump(path):
    """The parsed + resolved scene file as canonical text (factory/node/parser.cpp; host only)."""
    L = lib()
    need = c_uint64(0)
    _check(L.btrc_scene_file_dump(path.encode(), None, 0, byref(need)))
    buf = ctypes.create_string_buffer(need.value)
    _check(L.btrc_scene_file_dump(path.encode(), buf, need.value, byref(need)))
    return buf.value.decode()


PARSE_KINDS = {"int": 0, "float": 1, "bool": 2, "vec3": 3, "spectrum": 4, "degree": 5, "radian": 6, "transform3d": 7, "transform2d": 8}


def scene_file_parse(path, node_path, kind):
    """Node::parse<T> of the node at node_path (factory/node/node.cpp:276-325; host only). Returns a float32 array."""
    out = np.zeros(16, np.float32)
    n = c_int32(0)
    _check(lib().btrc_scene_file_parse(path.encode(), node_path.encode(), PARSE_KINDS[kind], out.ctypes.data, byref(n)))
    return out[:n.value].copy()


def render_file(path, **kw):
    """cli/main.cpp:17-81 as one call: load, render, run the scene's post-processors. Returns (color, stats)."""
    s = Scene(path, **kw)
    color, albedo, normal, stats = s.render()
    s.run_post_processors(color, albedo, normal)
    s.close()
    return color, stats
