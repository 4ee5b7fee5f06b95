"""ORACLE / reference build (test infrastructure only): ctypes face of oracle/_ref/libbtrcrefcpu.so.

That library is the REFERENCE'S OWN wavefront path tracer compiled for the host (oracle/refcpu/refcpu.cpp says exactly which
parts are reference code and which three are stand-ins: OptiX -> the reference's vendored lib/bvh, the texture unit, the shadow
miss program's OptiX getters).  Its stage calls take and return the arrays of oracle.front.OracleScene (same names, same
uncompacted convention), so tests compare oracle, reference and CUDA on identical batches.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

from .front import new_state, state_ptrs

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libbtrcrefcpu.so")
f32 = np.float32

_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(POINTER(ty))


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB)
        L.refcpu_open.restype = c_void_p
        L.refcpu_open.argtypes = [c_char_p]
        L.refcpu_last_error.restype = c_char_p
        L.refcpu_close.argtypes = [c_void_p]
        L.refcpu_jit_seconds.restype = c_double
        L.refcpu_jit_seconds.argtypes = [c_void_p]
        L.refcpu_world_diagonal.restype = c_float
        L.refcpu_world_diagonal.argtypes = [c_void_p]
        L.refcpu_kernel_source.restype = c_size_t
        L.refcpu_kernel_source.argtypes = [c_void_p, c_char_p, c_size_t]
        for name in ("refcpu_info", "refcpu_get_triangles", "refcpu_get_geometry_tables", "refcpu_closest_hit", "refcpu_generate", "refcpu_shade",
                     "refcpu_medium", "refcpu_shadow", "refcpu_render", "refcpu_set_spp"):
            getattr(L, name).argtypes = None
        _lib = L
    return _lib


class ReferenceScene:
    """A scene file loaded, committed and compiled by the reference's own code (cli/main.cpp:17-71)."""

    def __init__(self, filename: str):
        self.L = lib()
        self.h = self.L.refcpu_open(os.path.abspath(filename).encode())
        if not self.h:
            raise RuntimeError("reference failed on %s: %s" % (filename, self.L.refcpu_last_error().decode()))
        self.h = c_void_p(self.h)
        info = np.zeros(9, np.int64)
        self.L.refcpu_info(self.h, _ptr(info, c_int64))
        self.width, self.height, self.spp, self.state_count = (int(v) for v in info[:4])
        self._has_medium = bool(info[4])
        self.ntri, self.ninst, self._vol_prim_medium_id = int(info[5]), int(info[6]), int(info[7])
        # core/scene.cpp:136-161: instances without a medium get the id `medium_indices.size() - 1`, evaluated after the nullptr key
        # exists: one PAST the table (access_medium's `unreachable` default).  Oracle and product give such rays the
        # volume-primitive medium, the id just below it.
        self.null_medium_id = int(info[8])

    def close(self):
        if self.h:
            self.L.refcpu_close(self.h)
            self.h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.refcpu_last_error().decode())

    def to_oracle_ids(self, o_med):
        """Reference medium ids -> the oracle's convention, in place on an (n, 4) o_med array (see null_medium_id)."""
        w = o_med[:, 3].view(np.uint32)
        w[w == self.null_medium_id] = self._vol_prim_medium_id
        return o_med

    def has_medium(self): return self._has_medium
    def vol_prim_medium_id(self): return self._vol_prim_medium_id
    def jit_seconds(self): return float(self.L.refcpu_jit_seconds(self.h))
    def world_diagonal(self): return float(self.L.refcpu_world_diagonal(self.h))
    def num_triangles(self): return self.ntri

    def kernel_source(self) -> str:
        n = self.L.refcpu_kernel_source(self.h, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        self.L.refcpu_kernel_source(self.h, buf, n + 1)
        return buf.value.decode()

    def triangles(self):
        verts = np.zeros((self.ntri, 9), f32)
        ip = np.zeros((self.ntri, 2), np.uint32)
        self.L.refcpu_get_triangles(self.h, _ptr(verts, c_float), _ptr(ip, c_uint32))
        return verts, ip

    def geometry_tables(self, inst, ntri):
        out = np.zeros((ntri, 6, 4), f32)
        self.L.refcpu_get_geometry_tables(self.h, c_int(inst), c_int(ntri), _ptr(out, c_float))
        return out

    def closest_hit(self, o_med, d_t1):
        o_med = np.ascontiguousarray(o_med, f32); d_t1 = np.ascontiguousarray(d_t1, f32)
        n = o_med.shape[0]
        flag = np.zeros(n, np.uint32); detail = np.zeros((n, 4), np.uint32)
        self.L.refcpu_closest_hit(self.h, c_int(n), _ptr(o_med, c_float), _ptr(d_t1, c_float), _ptr(flag, c_uint32), _ptr(detail, c_uint32))
        return flag, detail

    def generate(self, first, n):
        st = new_state(n)
        self._check(self.L.refcpu_generate(self.h, c_int64(first), c_int(n), *state_ptrs(st)))
        return st

    def shade(self, st, flag, detail):
        n = st["o_med"].shape[0]
        st = {k: np.ascontiguousarray(v.copy()) for k, v in st.items()}
        nxt = new_state(n)
        out = dict(status=np.zeros(n, np.uint32), next=nxt,
                   sh_o_med=np.zeros((n, 4), f32), sh_d_t1=np.zeros((n, 4), f32), sh_beta_li=np.zeros((n, 4), f32),
                   splat=np.zeros((n, 3), f32), albedo=np.zeros((n, 3), f32), normal=np.zeros((n, 3), f32))
        flag = np.ascontiguousarray(flag.copy(), np.uint32); detail = np.ascontiguousarray(detail, np.uint32)
        self._check(self.L.refcpu_shade(self.h, c_int(n), *state_ptrs(st), _ptr(flag, c_uint32), _ptr(detail, c_uint32), _ptr(out["status"], c_uint32),
                                        *state_ptrs(nxt), _ptr(out["sh_o_med"], c_float), _ptr(out["sh_d_t1"], c_float), _ptr(out["sh_beta_li"], c_float),
                                        _ptr(out["splat"], c_float), _ptr(out["albedo"], c_float), _ptr(out["normal"], c_float)))
        return out

    def medium(self, st, flag, detail):
        n = st["o_med"].shape[0]
        st = {k: np.ascontiguousarray(v.copy()) for k, v in st.items()}
        flag = np.ascontiguousarray(flag.copy(), np.uint32); detail = np.ascontiguousarray(detail, np.uint32)
        nxt = new_state(n)
        out = dict(status=np.zeros(n, np.uint32), next=nxt, sh_o_med=np.zeros((n, 4), f32), sh_d_t1=np.zeros((n, 4), f32),
                   sh_beta_li=np.zeros((n, 4), f32), splat=np.zeros((n, 3), f32))
        self._check(self.L.refcpu_medium(self.h, c_int(n), *state_ptrs(st), _ptr(flag, c_uint32), _ptr(detail, c_uint32), _ptr(out["status"], c_uint32),
                                         *state_ptrs(nxt), _ptr(out["sh_o_med"], c_float), _ptr(out["sh_d_t1"], c_float), _ptr(out["sh_beta_li"], c_float),
                                         _ptr(out["splat"], c_float)))
        return st, flag, out

    def shadow(self, sh_o_med, sh_d_t1, sh_beta_li, pixel):
        """Returns (occluded u8[n], contribution f32[n,3]): what the film received for each shadow ray."""
        sh_o_med, sh_d_t1, sh_beta_li = (np.ascontiguousarray(a, f32) for a in (sh_o_med, sh_d_t1, sh_beta_li))
        pixel = np.ascontiguousarray(pixel, np.uint32)
        n = sh_o_med.shape[0]
        occ = np.zeros(n, np.uint8); contrib = np.zeros((n, 3), f32)
        self._check(self.L.refcpu_shadow(self.h, c_int(n), _ptr(sh_o_med, c_float), _ptr(sh_d_t1, c_float), _ptr(sh_beta_li, c_float),
                                         _ptr(pixel, c_uint32), _ptr(occ, c_uint8), _ptr(contrib, c_float)))
        return occ, contrib

    def set_spp(self, spp):
        self.L.refcpu_set_spp(self.h, c_int(spp))
        self.spp = spp

    def render(self, raw=False):
        color = np.zeros((self.height, self.width, 4), f32)
        albedo = np.zeros((self.height, self.width, 4), f32)
        normal = np.zeros((self.height, self.width, 4), f32)
        rr = np.zeros((self.height, self.width, 4), f32) if raw else None
        rw = np.zeros((self.height, self.width), f32) if raw else None
        stats = np.zeros(7, np.uint64)
        sec = c_double(0)
        self._check(self.L.refcpu_render(self.h, _ptr(color, c_float), _ptr(albedo, c_float), _ptr(normal, c_float), _ptr(rr, c_float), _ptr(rw, c_float),
                                         _ptr(stats, c_uint64), ctypes.byref(sec)))
        names = ["n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"]
        return dict(color=color, albedo=albedo, normal=normal, raw_radiance=rr, raw_weight=rw, seconds=sec.value,
                    stats={k: int(v) for k, v in zip(names, stats)})
