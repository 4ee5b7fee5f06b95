"""ORACLE front-end (test infrastructure only).

Python/numpy restatement of the reference's scene boundary -- JSON scene files, typed node parsing,
transforms and OBJ loading -- feeding the C++ oracle (oracle/liborc.so).  The product has its own
C++ loader (btrc_b200/csrc/host); this file exists so that the two can be checked against each other.

Reference files followed (relative to /root/reference/src):
  factory/btrc/factory/node/parser.cpp:16-45,59-76,106-252   JSON -> node tree, $include / $reference,
                                                          the std::to_string(float) round trip
  factory/btrc/factory/node/node.cpp:50-214              Vec3f / Spectrum / angle / Transform3D parsing
  common/btrc/utils/math/mat4.cpp:108-199, transform.cpp:92-130   matrices (incl. the rotate_z quirk)
  common/btrc/utils/triangle_mesh_loader.cpp:10-149 + lib/tinyobjloader (triangulate=true; quad rule at
                                                          tiny_obj_loader.h:1429-1520)
  factory/btrc/factory/scene.cpp:5-94, cli/main.cpp:36-66 entities / camera / renderer
"""
from __future__ import annotations

import ctypes
import json
import os
import re
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libm = ctypes.CDLL("libm.so.6")
for _n in ("sinf", "cosf"):
    getattr(_libm, _n).restype = c_float
    getattr(_libm, _n).argtypes = [c_float]

f32 = np.float32
PI = f32(3.1415926535)


# --------------------------------------------------------------------------- library loading
_lib = None
_ref = None


def _ptr(a, ty):
    return a.ctypes.data_as(POINTER(ty)) if a is not None else None


def lib():
    """Load oracle/liborc.so (build it with `make -C oracle`)."""
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liborc.so missing: run `make -C oracle`")
        L = ctypes.CDLL(path)
        L.orc_scene_new.restype = c_void_p
        L.orc_scene_free.argtypes = [c_void_p]
        L.orc_add_texture_constant.argtypes = [c_void_p, c_float, c_float, c_float]
        L.orc_add_material.argtypes = [c_void_p, c_int, POINTER(c_int32), c_int, c_int]
        L.orc_add_image.argtypes = [c_void_p, c_int, c_int, c_int, c_int, POINTER(c_float), c_int, c_int, c_int, c_int, c_int, POINTER(c_float)]
        L.orc_add_texture_array.argtypes = [c_void_p, c_int]
        L.orc_set_image_border_levels.argtypes = [c_void_p, c_int, c_int]
        L.orc_set_image_minmax.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_float)]
        L.orc_render_pixels.argtypes = [c_void_p, c_int, POINTER(c_uint32), c_int, c_int, POINTER(c_float), POINTER(c_uint64)]
        L.orc_add_texture_binary.argtypes = [c_void_p, c_int, c_int, c_int]
        L.orc_add_texture_transform.argtypes = [c_void_p, c_int, POINTER(c_float)]
        L.orc_eval_texture2d.argtypes = [c_void_p, c_int, c_int, POINTER(c_float), POINTER(c_float)]
        L.orc_add_texture3d_constant.argtypes = [c_void_p, c_float, c_float, c_float]
        L.orc_add_texture3d_array.argtypes = [c_void_p, c_int]
        L.orc_add_texture3d_binary.argtypes = [c_void_p, c_int, c_int, c_int]
        L.orc_texture3d_extremum.argtypes = [c_void_p, c_int, c_int]
        L.orc_texture3d_extremum.restype = c_float
        L.orc_add_medium_homogeneous.argtypes = [c_void_p, c_float, POINTER(c_float), c_float]
        L.orc_add_medium_heterogeneous.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_float)]
        L.orc_add_volume.argtypes = [c_void_p] + [POINTER(c_float)] * 4 + [c_int, c_int]
        L.orc_set_gradient_sky.argtypes = [c_void_p, POINTER(c_float), POINTER(c_float), POINTER(c_float)]
        L.orc_set_ibl.argtypes = [c_void_p, c_int, POINTER(c_float), c_int, c_int]
        L.orc_get_ibl_table.argtypes = [c_void_p, POINTER(c_float)]
        L.orc_medium.argtypes = ([c_void_p, c_int] + [POINTER(c_float)] * 2 + [POINTER(c_uint32)] + [POINTER(c_float)] * 2 +
                                 [POINTER(c_uint64), POINTER(c_float), POINTER(c_uint32), POINTER(c_uint32)] + [POINTER(c_uint32)] +
                                 [POINTER(c_float)] * 2 + [POINTER(c_uint32)] + [POINTER(c_float)] * 2 + [POINTER(c_uint64), POINTER(c_float)] +
                                 [POINTER(c_float)] * 4)
        L.orc_shadow_tr.argtypes = [c_void_p, c_int] + [POINTER(c_float)] * 3
        L.orc_has_medium.argtypes = [c_void_p]
        L.orc_bsdf_probe.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_float), c_int] + [POINTER(c_float)] * 8
        L.orc_add_mesh.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_float), POINTER(c_uint8), POINTER(c_float), POINTER(c_float)]
        L.orc_add_instance.argtypes = [c_void_p, c_int, c_int, POINTER(c_float), POINTER(c_float), c_int, POINTER(c_float), c_int, c_int]
        L.orc_set_camera.argtypes = [c_void_p, POINTER(c_float), POINTER(c_float), POINTER(c_float), c_float]
        L.orc_set_film.argtypes = [c_void_p, c_int, c_int]
        L.orc_set_renderer.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_int, c_float, c_float]
        L.orc_commit.argtypes = [c_void_p]
        L.orc_set_sampler.argtypes = [c_void_p, c_int]
        L.orc_owen_scrambled_radical_inverse.argtypes = [c_int, c_uint64, c_uint64]
        L.orc_owen_scrambled_radical_inverse.restype = c_float
        L.orc_prime.argtypes = [c_int]
        L.orc_prime.restype = c_int
        L.orc_num_triangles.argtypes = [c_void_p]
        L.orc_vol_prim_medium_id.argtypes = [c_void_p]
        L.orc_vol_prim_medium_id.restype = c_uint32
        L.orc_world_diagonal.argtypes = [c_void_p]
        L.orc_world_diagonal.restype = c_float
        L.orc_get_triangles.argtypes = [c_void_p, POINTER(c_float), POINTER(c_uint32)]
        L.orc_get_mesh_tables.argtypes = [c_void_p, c_int, POINTER(c_float)]
        L.orc_closest_hit.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_float), POINTER(c_uint32), POINTER(c_uint32), c_int]
        L.orc_any_hit.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_float), POINTER(c_uint8)]
        L.orc_generate.argtypes = [c_void_p, c_int64, c_int] + [POINTER(c_float)] * 2 + [POINTER(c_uint32)] + [POINTER(c_float)] * 2 + [POINTER(c_uint64), POINTER(c_float)]
        L.orc_shade.argtypes = ([c_void_p, c_int] + [POINTER(c_float)] * 2 + [POINTER(c_uint32)] + [POINTER(c_float)] * 2 +
                                [POINTER(c_uint64), POINTER(c_float), POINTER(c_uint32), POINTER(c_uint32)] + [POINTER(c_uint32)] +
                                [POINTER(c_float)] * 2 + [POINTER(c_uint32)] + [POINTER(c_float)] * 2 + [POINTER(c_uint64), POINTER(c_float)] +
                                [POINTER(c_float)] * 3 + [POINTER(c_float)] * 3)
        L.orc_render.argtypes = [c_void_p, c_int, c_int] + [POINTER(c_float)] * 5 + [POINTER(c_uint64)]
        L.orc_murmur_pixel.argtypes = [c_uint32, c_uint32]
        L.orc_murmur_pixel.restype = c_uint64
        L.orc_pcg_stream.argtypes = [c_uint32, c_uint32, c_int, c_int, POINTER(c_uint32), POINTER(c_float), POINTER(c_uint64)]
        L.orc_alias_build.argtypes = [c_int, POINTER(c_float), POINTER(c_float), POINTER(c_uint32)]
        L.orc_intersection_offset.argtypes = [c_int] + [POINTER(c_float)] * 4
        L.orc_num_threads.restype = c_int
        _lib = L
    return _lib


def ref_lib():
    """Load oracle/_ref/libbvhref.so: the reference's own lib/bvh (built here by `make -C oracle ref`)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libbvhref.so")
        if not os.path.exists(path):
            return None
        R = ctypes.CDLL(path)
        R.ref_bvh_build.restype = c_void_p
        R.ref_bvh_build.argtypes = [c_int, POINTER(c_float), c_int]
        R.ref_bvh_free.argtypes = [c_void_p]
        R.ref_bvh_closest.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_float), POINTER(c_int32), POINTER(c_float), POINTER(c_uint64)]
        R.ref_bvh_render_normals.argtypes = [c_void_p] + [POINTER(c_float)] * 3 + [c_float, c_int, c_int] + [POINTER(c_float)] * 3
        _ref = R
    return _ref


# --------------------------------------------------------------------------- JSON node tree
def _strip_comments(src: str) -> str:
    """nlohmann parse(..., ignore_comments=true): // and /* */ outside strings."""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == '"':
            j = i + 1
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1])
            i = j + 1
        elif src.startswith("//", i):
            while i < n and src[i] != "\n":
                i += 1
        elif src.startswith("/*", i):
            i = src.index("*/", i) + 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


def _to_value_string(v):
    """json_to_node (parser.cpp:33-43): every leaf becomes a string."""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        return "%f" % float(f32(v))  # std::to_string(json.get<float>())
    if isinstance(v, int):
        return str(int(np.int32(v)))
    return str(v)


class Node:
    """Group (dict) / Array (list) / Value (str) wrapper with the typed parsers of node.cpp."""

    def __init__(self, raw):
        self.raw = raw

    # -- structure
    def is_group(self): return isinstance(self.raw, dict)
    def is_array(self): return isinstance(self.raw, list)
    def is_value(self): return isinstance(self.raw, str)

    def find(self, name):
        if not self.is_group():
            raise ValueError("group node expected")
        return Node(self.raw[name]) if name in self.raw else None

    def child(self, name):
        r = self.find(name)
        if r is None:
            raise KeyError(f"child {name} not found")
        return r

    def elements(self):
        return [Node(e) for e in self.raw]

    # -- values
    def as_float(self):
        if not self.is_value():
            raise ValueError("value node expected")
        return f32(float(self.raw))  # std::stof

    def as_int(self):
        return int(float(self.raw)) if "." in self.raw else int(self.raw)  # std::stoi stops at '.'

    def as_bool(self): return self.raw == "true"
    def as_str(self): return self.raw

    def float_or(self, name, d):
        c = self.find(name)
        return f32(d) if c is None else c.as_float()

    def int_or(self, name, d):
        c = self.find(name)
        return d if c is None else c.as_int()

    def bool_or(self, name, d):
        c = self.find(name)
        return d if c is None else c.as_bool()

    def as_vec3(self):  # node.cpp:50-74 (Spectrum: :100-124 is identical)
        if self.is_array():
            e = self.elements()
            if len(e) == 1:
                v = e[0].as_float()
                return np.array([v, v, v], f32)
            if len(e) == 3:
                return np.array([e[0].as_float(), e[1].as_float(), e[2].as_float()], f32)
            raise ValueError(f"unexpected array size: {len(e)}")
        if self.is_value():
            v = self.as_float()
            return np.array([v, v, v], f32)
        raise ValueError("unexpected group node")

    def as_degree(self):  # node.cpp:76-86
        c = self.find("deg")
        if c is not None:
            return c.as_float()
        c = self.find("rad")
        if c is not None:
            return f32(f32(180.0) / PI) * c.as_float()
        raise ValueError("invalid angle node")

    def as_radian(self):  # node.cpp:88-98
        c = self.find("deg")
        if c is not None:
            return f32(PI / f32(180.0)) * c.as_float()
        c = self.find("rad")
        if c is not None:
            return c.as_float()
        raise ValueError("invalid angle node")

    def as_transform3d(self):  # node.cpp:166-213
        if self.is_array():
            ret = Transform3D()
            for e in self.elements():
                ret = e.as_transform3d().mul(ret)
            return ret
        c = self.find("translate")
        if c is not None:
            t = c.as_vec3()
            return Transform3D.translate(t[0], t[1], t[2])
        c = self.find("rotate")
        if c is not None:
            return Transform3D.rotate(c.child("axis").as_vec3(), c.child("angle").as_radian())
        for k, fn in (("rotate_x", Transform3D.rotate_x), ("rotate_y", Transform3D.rotate_y), ("rotate_z", Transform3D.rotate_z)):
            c = self.find(k)
            if c is not None:
                return fn(c.as_radian())
        c = self.find("scale")
        if c is not None:
            s = c.as_vec3()
            return Transform3D.scale(s[0], s[1], s[2])
        raise ValueError("unknown transform3d type")


def _node_transform2d(node):
    """node.cpp:126-164 -> 3x3 row-major matrix (f32).  translate / scale need parse<Vec2f>, which the reference never
    instantiates (node.cpp:321-324: "unimplemented"); the 3x3 product is the intended one (Mat3 operator* in
    utils/math/mat3.cpp:108-123 reads a fourth column that does not exist)."""
    if node.is_array():
        ret = np.eye(3, dtype=f32)
        for e in node.elements():
            ret = (_node_transform2d(e).astype(f32) @ ret).astype(f32)
        return ret
    if node.find("translate") is not None or node.find("scale") is not None:
        raise ValueError("unimplemented")
    c = node.find("rotate")
    if c is not None:
        rad = c.child("angle").as_radian()
        S, C = _sin(rad), _cos(rad)
        return np.array([[C, -S, 0], [S, C, 0], [0, 0, 1]], f32)
    c = node.find("inverse_u")
    if c is not None and c.as_bool():
        return np.array([[-1, 0, 1], [0, 1, 0], [0, 0, 0]], f32)
    c = node.find("inverse_v")
    if c is not None and c.as_bool():
        return np.array([[1, 0, 0], [0, -1, 1], [0, 0, 1]], f32)
    c = node.find("swap_uv")
    if c is not None and c.as_bool():
        return np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]], f32)
    raise ValueError("unknown transform2d type")


def _json_to_tree(j):
    if isinstance(j, dict):
        return {k: _json_to_tree(v) for k, v in j.items()}
    if isinstance(j, list):
        return [_json_to_tree(v) for v in j]
    return _to_value_string(j)


def _resolve_refs(path, node):
    """parser.cpp:135-252.  `path` = list of ancestor containers; returns the replacement for `node`."""
    if isinstance(node, dict):
        path.append(node)
        for k in list(node.keys()):
            node[k] = _resolve_refs(path, node[k])
        path.pop()
        return node
    if isinstance(node, list):
        path.append(node)
        for i in range(len(node)):
            node[i] = _resolve_refs(path, node[i])
        path.pop()
        return node
    if node.startswith("$reference{"):
        if not node.endswith("}"):
            raise ValueError("'}' for '$reference{' is not found")
        return _find_node(path, node[11:-1])
    return node


def _find_node(cur_path, p):
    secs = p.split("/")
    if not secs or secs == [""]:
        raise ValueError("invalid node path: " + p)
    start = secs[0]
    if start == "$root":
        si = 0
    else:
        si = len(cur_path) - 1
        while si >= 0:
            n = cur_path[si]
            if isinstance(n, dict):
                if start in n:
                    break
            elif isinstance(n, list):
                if start.isdigit() and int(start) < len(n):
                    break
            si -= 1
    if si < 0:
        raise ValueError("invalid node path: " + p)
    node = cur_path[si]
    for i, sec in enumerate(secs):
        if i == 0 and sec == "$root":
            continue
        if isinstance(node, dict):
            if sec not in node:
                raise ValueError("invalid node path: " + p)
            node = node[sec]
        elif isinstance(node, list):
            if not sec.isdigit() or int(sec) >= len(node):
                raise ValueError("invalid node path: " + p)
            node = node[int(sec)]
        else:
            raise ValueError("invalid node path: " + p)
    if isinstance(node, str) and "$reference{" in node:
        node = _find_node(cur_path, node[11:-1])
    return node


def parse_scene_file(filename: str) -> Node:
    scene_dir = os.path.dirname(os.path.abspath(filename))
    src = open(filename).read()
    while True:  # process_includes (parser.cpp:106-133)
        beg = src.find('"$include{')
        if beg < 0:
            break
        end = src.index('}"', beg + 10)
        inc = src[beg + 10:end]
        inc_path = inc if os.path.isabs(inc) else os.path.join(scene_dir, inc)
        src = src[:beg] + open(inc_path).read() + src[end + 2:]
    tree = _json_to_tree(json.loads(_strip_comments(src)))
    if not isinstance(tree, dict):
        raise ValueError("root node is not a group")
    path = [tree]
    for k in list(tree.keys()):
        tree[k] = _resolve_refs(path, tree[k])
    return Node(tree)


# --------------------------------------------------------------------------- transforms
def _m4(rows):
    return np.array(rows, dtype=f32).reshape(4, 4)


def _matmul(a, b):  # mat4.cpp:186-199 (left-to-right sum in fp32)
    r = np.zeros((4, 4), f32)
    for i in range(4):
        for c in range(4):
            r[i, c] = f32(f32(f32(a[i, 0] * b[0, c]) + f32(a[i, 1] * b[1, c])) + f32(a[i, 2] * b[2, c])) + f32(a[i, 3] * b[3, c])
    return r


def _sin(x): return f32(_libm.sinf(c_float(float(x))))
def _cos(x): return f32(_libm.cosf(c_float(float(x))))


def _rot_axis(axis, rad):  # mat4.cpp:117-145
    ln = np.sqrt(f32(f32(f32(axis[0] * axis[0]) + f32(axis[1] * axis[1])) + f32(axis[2] * axis[2])))
    inv = f32(1) / f32(ln)
    a = np.array([inv * axis[0], inv * axis[1], inv * axis[2]], f32)
    s, c = _sin(rad), _cos(rad)
    one = f32(1)
    x, y, z = a
    m = np.zeros((4, 4), f32)
    m[0, 0] = f32(x * x) + f32(f32(one - f32(x * x)) * c)
    m[0, 1] = f32(f32(x * y) * f32(one - c)) - f32(z * s)
    m[0, 2] = f32(f32(x * z) * f32(one - c)) + f32(y * s)
    m[1, 0] = f32(f32(x * y) * f32(one - c)) + f32(z * s)
    m[1, 1] = f32(y * y) + f32(f32(one - f32(y * y)) * c)
    m[1, 2] = f32(f32(y * z) * f32(one - c)) - f32(x * s)
    m[2, 0] = f32(f32(x * z) * f32(one - c)) - f32(y * s)
    m[2, 1] = f32(f32(y * z) * f32(one - c)) + f32(x * s)
    m[2, 2] = f32(z * z) + f32(f32(one - f32(z * z)) * c)
    m[3, 3] = 1
    # Mat4 is column-major (at(r, c) = data[c][r], mat4.cpp:98-106) and Mat4::rotate writes data[i][j] with the textbook ROW-major
    # entries: the matrix the reference uses is the transpose (a turn by -rad).  Pinned by tests/test_reference_host_pins_cpu.py.
    return np.ascontiguousarray(m.T)


def _rot_x(rad):
    s, c = _sin(rad), _cos(rad)
    return _m4([1, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, 1])


def _rot_y(rad):
    s, c = _sin(rad), _cos(rad)
    return _m4([c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1])


def _rot_z(rad):
    s, c = _sin(rad), _cos(rad)
    return _m4([c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1])


class Transform3D:
    def __init__(self, mat=None, inv=None):
        self.mat = np.eye(4, dtype=f32) if mat is None else mat
        self.inv = np.eye(4, dtype=f32) if inv is None else inv

    def mul(self, b):  # transform.cpp:127-130: (a.mat * b.mat, b.inv * a.inv)
        return Transform3D(_matmul(self.mat, b.mat), _matmul(b.inv, self.inv))

    @staticmethod
    def translate(x, y, z):
        return Transform3D(_m4([1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z, 0, 0, 0, 1]),
                           _m4([1, 0, 0, -x, 0, 1, 0, -y, 0, 0, 1, -z, 0, 0, 0, 1]))

    @staticmethod
    def rotate(axis, rad): return Transform3D(_rot_axis(axis, rad), _rot_axis(axis, -rad))
    @staticmethod
    def rotate_x(rad): return Transform3D(_rot_x(rad), _rot_x(-rad))
    @staticmethod
    def rotate_y(rad): return Transform3D(_rot_y(rad), _rot_y(-rad))

    @staticmethod
    def rotate_z(rad):
        # transform.cpp:112-115: the reference passes -rad for BOTH mat and inv (quirk kept)
        return Transform3D(_rot_z(-rad), _rot_z(-rad))

    @staticmethod
    def scale(x, y, z):
        one = f32(1)
        return Transform3D(_m4([x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1]),
                           _m4([one / x, 0, 0, 0, 0, one / y, 0, 0, 0, 0, one / z, 0, 0, 0, 0, 1]))


# --------------------------------------------------------------------------- OBJ
_OBJ_CACHE = {}
_OBJ_VERTICES = {}  # every "v" of the file, used or not: what transform_to_unit_cube takes its bounding box from


def load_obj(filename: str):
    """Cached by (path, mtime): scenes instance one OBJ many times and every entity creates its own geometry object."""
    key = (filename, os.path.getmtime(filename))
    if key not in _OBJ_CACHE:
        _OBJ_CACHE[key] = _load_obj(filename)
    return _OBJ_CACHE[key]


def _ear_clip(face, vs):
    """tinyobj's built-in ear clipping for polygons with more than four corners (tiny_obj_loader.h:1521-1813), float arithmetic
    in its order: projection axes from the first non-degenerate corner, candidate ears from a moving start, accepted when
    turn * "area" >= 0 and no other remaining corner is inside (pnpoly, :1387-1398); gives up after a round without progress."""
    n = len(face)
    P = [np.array(vs[c[0]], f32) for c in face]
    ax0, ax1 = 1, 2
    eps = f32(np.finfo(f32).eps)
    for k in range(n):
        a, b, c = P[k], P[(k + 1) % n], P[(k + 2) % n]
        e0, e1 = (b - a).astype(f32), (c - b).astype(f32)
        cx = abs(f32(f32(e0[1] * e1[2]) - f32(e0[2] * e1[1])))
        cy = abs(f32(f32(e0[2] * e1[0]) - f32(e0[0] * e1[2])))
        cz = abs(f32(f32(e0[0] * e1[1]) - f32(e0[1] * e1[0])))
        if cx > eps or cy > eps or cz > eps:
            if not (cx > cy and cx > cz):
                ax0 = 0
                if cz > cx and cz > cy:
                    ax1 = 1
            break

    def inside(vx, vy, tx, ty):
        c = False
        j = 2
        for i in range(3):
            if (vy[i] > ty) != (vy[j] > ty):
                if tx < f32(f32(f32(f32(vx[j] - vx[i]) * f32(ty - vy[i])) / f32(vy[j] - vy[i])) + vx[i]):
                    c = not c
            j = i
        return c

    rem = list(range(n))
    out = []
    start, budget, last = 0, n, n
    while len(rem) > 3 and budget > 0:
        m = len(rem)
        if start >= m:
            start -= m
        if last != m:
            last, budget = m, m
        else:
            budget -= 1
        tri = [rem[(start + k) % m] for k in range(3)]
        vx = [f32(P[t][ax0]) for t in tri]
        vy = [f32(P[t][ax1]) for t in tri]
        e0x, e0y, e1x, e1y = f32(vx[1] - vx[0]), f32(vy[1] - vy[0]), f32(vx[2] - vx[1]), f32(vy[2] - vy[1])
        turn = f32(f32(e0x * e1y) - f32(e0y * e1x))
        area = f32(f32(f32(vx[0] * vy[1]) - f32(vy[0] * vx[1])) * f32(0.5))
        if f32(turn * area) < 0:
            start += 1
            continue
        if any(inside(vx, vy, f32(P[rem[(start + o) % m]][ax0]), f32(P[rem[(start + o) % m]][ax1])) for o in range(3, m)):
            start += 1
            continue
        out += [face[t] for t in tri]
        del rem[(start + 1) % m]
    if len(rem) == 3:
        out += [face[t] for t in rem]
    return out


def _load_obj(filename: str):
    """tinyobj (triangulate=true) + TriangleMeshLoader: returns per-corner arrays in face order.

    pos (ntri*3, 3) f32, nor (ntri*3, 3) f32, has_nor (ntri*3) u8, uv (ntri*3, 2) f32.
    Triangles stay as they are, quads split along the shorter diagonal (tiny_obj_loader.h:1429-1520);
    larger polygons go through tinyobj's ear clipping (_ear_clip).
    """
    vs, vns, vts = [], [], []
    corners = []  # (vi, ti, ni)

    def fix(idx, n):
        return idx - 1 if idx > 0 else n + idx

    with open(filename) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                vs.append([f32(float(t)) for t in tok[1:4]])
            elif tok[0] == "vn":
                vns.append([f32(float(t)) for t in tok[1:4]])
            elif tok[0] == "vt":
                vts.append(([f32(float(t)) for t in tok[1:3]] + [f32(0), f32(0)])[:2])  # a missing v is 0 (tinyobj parseReal2 default)
            elif tok[0] == "f":
                face = []
                for t in tok[1:]:
                    p = t.split("/")
                    vi = fix(int(p[0]), len(vs))
                    ti = fix(int(p[1]), len(vts)) if len(p) > 1 and p[1] else -1
                    ni = fix(int(p[2]), len(vns)) if len(p) > 2 and p[2] else -1
                    face.append((vi, ti, ni))
                if len(face) == 3:
                    corners += face
                elif len(face) == 4:
                    v0, v1, v2, v3 = (np.array(vs[c[0]], f32) for c in face)
                    e02, e13 = v2 - v0, v3 - v1
                    s02 = f32(f32(f32(e02[0] * e02[0]) + f32(e02[1] * e02[1])) + f32(e02[2] * e02[2]))
                    s13 = f32(f32(f32(e13[0] * e13[0]) + f32(e13[1] * e13[1])) + f32(e13[2] * e13[2]))
                    if s02 < s13:
                        corners += [face[0], face[1], face[2], face[0], face[2], face[3]]
                    else:
                        corners += [face[0], face[1], face[3], face[1], face[2], face[3]]
                elif len(face) > 4:
                    corners += _ear_clip(face, vs)
    n = len(corners)
    pos = np.zeros((n, 3), f32)
    nor = np.zeros((n, 3), f32)
    has_nor = np.zeros(n, np.uint8)
    uv = np.zeros((n, 2), f32)
    for i, (vi, ti, ni) in enumerate(corners):
        pos[i] = vs[vi]
        if ni >= 0:
            nor[i] = vns[ni]
            has_nor[i] = 1
        if ti >= 0:
            uv[i] = vts[ti]
    _OBJ_VERTICES[filename] = np.array(vs, f32).reshape(-1, 3)
    return pos, nor, has_nor, uv


# --------------------------------------------------------------------------- scene
class OracleScene:
    """Builds an oracle scene from a Btrc JSON scene file (the same file the product loads)."""

    MAT_TYPES = {"diffuse": 0, "black": 1, "disney": 2, "metal": 3, "glass": 4, "mirror": 5, "invisible": 6}
    SLOTS = ["color", "metallic", "roughness", "transmission", "ior", "specular", "specular_tint", "anisotropic", "sheen", "sheen_tint",
             "clearcoat", "clearcoat_gloss", "transmission_roughness"]
    ADDR = {"clamp": 0, "wrap": 1, "mirror": 2, "border": 3}

    def __init__(self, filename: str | None = None):
        self.L = lib()
        self.h = self.L.orc_scene_new()
        self.width = self.height = 0
        self.spp = 1
        self.meshes = {}      # filename -> mesh id
        self.mesh_data = []   # (pos, nor, has_nor, uv) per mesh id
        self.instances = []   # dicts
        self.want_albedo = self.want_normal = False
        self.scene_dir = "."
        self._pool = {}       # (kind, id(raw node)) -> object id  (Context::create object pooling, context.h:205-236)
        self.media = []       # (priority, creation order, builder) resolved at commit
        if filename:
            self._load(filename)

    def __del__(self):
        try:
            self.L.orc_scene_free(self.h)
        except Exception:
            pass

    def resolve(self, p):  # factory/path_resolver.cpp:6-17
        return os.path.normpath(os.path.abspath(p.replace("$env{scene_directory}", self.scene_dir)))

    # -- construction helpers (usable without a JSON file, e.g. the lib/bvh golden test)
    def add_mesh_arrays(self, pos, nor=None, has_nor=None, uv=None, frame_pos=None):
        """frame_pos: the positions as read from the file when `pos` went through transform_to_unit_cube (the reference derives the
        geometry frames before that rescale: triangle_mesh_loader.cpp:80-123 vs 193-213)."""
        pos = np.ascontiguousarray(pos, f32)
        frame_pos = None if frame_pos is None else np.ascontiguousarray(frame_pos, f32)
        ntri = pos.shape[0] // 3
        nor = None if nor is None else np.ascontiguousarray(nor, f32)
        has_nor = None if has_nor is None else np.ascontiguousarray(has_nor, np.uint8)
        uv = None if uv is None else np.ascontiguousarray(uv, f32)
        mid = self.L.orc_add_mesh(self.h, ntri, _ptr(pos, c_float), _ptr(nor, c_float), _ptr(has_nor, c_uint8), _ptr(uv, c_float), _ptr(frame_pos, c_float))
        self.mesh_data.append((pos, nor, has_nor, uv))
        return mid

    def _texture_desc(self, node):  # builtin/texture2d/description.cpp:28-80
        au = self.ADDR[node.find("address_mode_u").as_str()] if node.find("address_mode_u") is not None else 0
        av = self.ADDR[node.find("address_mode_v").as_str()] if node.find("address_mode_v") is not None else 0
        aw = self.ADDR[node.find("address_mode_w").as_str()] if node.find("address_mode_w") is not None else 0
        filt = node.find("filter").as_str() if node.find("filter") is not None else "linear"
        if filt not in ("linear", "point"):
            raise ValueError(f"unknown filter mode: {filt}")
        border = np.zeros(4, f32)
        bn = node.find("border")
        if bn is not None:
            if bn.is_array():
                e = bn.elements()
                if len(e) == 1:
                    border[:] = e[0].as_float()
                elif len(e) == 3:
                    border[:3] = [x.as_float() for x in e]
                else:
                    raise ValueError("invalid border array size")
            else:
                border[:] = bn.as_float()
        return au, av, aw, 1 if filt == "linear" else 0, node.bool_or("srgb_to_linear", False), border

    def add_image_2d(self, filename, desc):
        """utils/image.cpp:299-335 + cuda/array.cpp:136-149: 1-channel 8-bit -> UNorm8x1, 3/4-channel -> UNorm8x4 (alpha 255),
        .hdr -> F32x4 (alpha 1)."""
        from PIL import Image as PILImage

        au, av, aw, linear, srgb, border = desc
        if filename.lower().endswith(".hdr"):
            import cv2

            img = cv2.imread(filename, cv2.IMREAD_UNCHANGED)[..., ::-1].astype(f32)
            data = np.concatenate([img, np.ones(img.shape[:2] + (1,), f32)], axis=2)
            ch, srgb = 4, False
        else:
            im = PILImage.open(filename)
            if im.mode in ("L", "1"):
                data = (np.asarray(im.convert("L"), np.uint8).astype(f32) / f32(255.0))[..., None]
                ch = 1
            else:
                rgba = np.asarray(im.convert("RGBA") if im.mode in ("RGBA", "LA") else im.convert("RGB"), np.uint8)
                if rgba.shape[2] == 3:
                    rgba = np.concatenate([rgba, np.full(rgba.shape[:2] + (1,), 255, np.uint8)], axis=2)
                data = rgba.astype(f32) / f32(255.0)
                ch = 4
        data = np.ascontiguousarray(data, f32)
        h, w = data.shape[:2]
        img = self.L.orc_add_image(self.h, w, h, 0, ch, _ptr(data, c_float), int(bool(srgb)), au, av, aw, linear, _ptr(border, c_float))
        if not filename.lower().endswith(".hdr"):
            self.L.orc_set_image_border_levels(self.h, img, 255)   # UNorm8 storage: the border colour is held in the texel format
        return img

    def add_texture(self, node: Node):
        key = ("tex2d", id(node.raw))
        if key in self._pool:
            return self._pool[key]
        self._pool[key] = t = self._add_texture(node)
        return t

    def _add_texture(self, node: Node):
        # Context::create<Texture2D> (factory/context.h:211-220): literal number / array => constant
        if node.is_value() or node.is_array():
            v = node.as_vec3()
            return self.L.orc_add_texture_constant(self.h, v[0], v[1], v[2])
        ty = node.child("type").as_str()
        if ty == "constant":
            v = node.child("value").as_vec3()
            return self.L.orc_add_texture_constant(self.h, v[0], v[1], v[2])
        if ty == "array":
            img = self.add_image_2d(self.resolve(node.child("filename").as_str()), self._texture_desc(node))
            return self.L.orc_add_texture_array(self.h, img)
        if ty in ("add", "sub", "mul", "div"):
            lhs, rhs = self.add_texture(node.child("lhs")), self.add_texture(node.child("rhs"))
            return self.L.orc_add_texture_binary(self.h, {"add": 2, "sub": 3, "mul": 4, "div": 5}[ty], lhs, rhs)
        if ty == "transform":
            m = _node_transform2d(node.child("transform"))
            inner = self.add_texture(node.child("tex"))
            m6 = np.ascontiguousarray(m[:2].reshape(-1), f32)
            return self.L.orc_add_texture_transform(self.h, inner, _ptr(m6, c_float))
        raise ValueError(f"unknown texture2d type: {ty}")

    def const_texture(self, v):
        return self.L.orc_add_texture_constant(self.h, c_float(v), c_float(v), c_float(v))

    def add_material(self, node: Node):
        ty = node.child("type").as_str()
        if ty not in self.MAT_TYPES:
            raise ValueError(f"unknown material type: {ty}")
        tex = np.full(13, -1, np.int32)

        def req(name): tex[self.SLOTS.index(name)] = self.add_texture(node.child(name))

        def opt(name, default):
            n = node.find(name)
            tex[self.SLOTS.index(name)] = self.add_texture(n) if n is not None else self.const_texture(default)

        normal, st = -1, True
        if ty == "diffuse":  # diffuse.cpp:36-48
            req("color")
        elif ty == "disney":  # disney.cpp:844-894
            req("color"); req("metallic"); req("roughness")
            opt("transmission", 0); opt("ior", 1.5); opt("specular", 1); opt("specular_tint", 0); opt("anisotropic", 0)
            opt("sheen", 0); opt("sheen_tint", 0); opt("clearcoat", 0); opt("clearcoat_gloss", 1)
            n = node.find("transmission_roughness")
            tex[12] = self.add_texture(n) if n is not None else tex[2]
        elif ty == "metal":  # metal.cpp:50-77
            req("color"); req("roughness"); opt("anisotropic", 0)
        elif ty == "glass":  # glass.cpp:141-157
            req("color"); req("ior")
        elif ty == "mirror":  # mirror.cpp:91-103
            req("color")
        if ty in ("diffuse", "disney", "metal"):
            st = node.bool_or("shadow_terminator_term", True)
        if ty in ("diffuse", "disney", "metal", "glass", "mirror") and node.find("normal") is not None:
            normal = self.add_texture(node.child("normal"))
        return self.L.orc_add_material(self.h, self.MAT_TYPES[ty], _ptr(tex, c_int32), normal, 1 if st else 0)

    # -- 3-D textures and media
    def add_texture3d(self, node: Node):
        key = ("tex3d", id(node.raw))
        if key in self._pool:
            return self._pool[key]
        if node.is_value() or node.is_array():
            v = node.as_vec3()
            t = self.L.orc_add_texture3d_constant(self.h, v[0], v[1], v[2])
        else:
            ty = node.child("type").as_str()
            if ty == "constant":
                v = node.child("value").as_vec3()
                t = self.L.orc_add_texture3d_constant(self.h, v[0], v[1], v[2])
            elif ty == "array":  # utils/cuda/array.cpp:313-385 (text grid)
                au, av, aw, linear, srgb, border = self._texture_desc(node)
                with open(self.resolve(node.child("text").as_str())) as f:
                    text = f.read()
                tok = text.split()
                comp, ch, w, h, d = tok[0], int(tok[1]), int(tok[2]), int(tok[3]), int(tok[4])
                if ch not in (1, 3, 4):
                    raise ValueError(f"unsupported channel count: {ch}")
                minmax = None
                if comp == "byte":
                    # array.cpp:331-355: `uint8_t comp; fin >> comp;` extracts ONE non-blank character per component; the grid is
                    # uploaded as SNorm8 (texel = max(int8 / 127, -1)), a 3-channel texel gets 255 as its fourth component; the
                    # extrema are the int8 extrema times float(1 / 127) (find_minmax_impl<int8_t, N, true>, array.cpp:436-510)
                    import re
                    head = re.match(r"\s*\S+\s+\S+\s+\S+\s+\S+\s+\S+", text)
                    chars = np.frombuffer("".join(text[head.end():].split()).encode("latin-1"), np.uint8)[:w * h * d * ch]
                    if chars.size < w * h * d * ch:
                        raise ValueError("failed to load byte grid")
                    i8 = chars.view(np.int8).reshape(d * h * w, ch)
                    if ch == 3:
                        i8 = np.concatenate([i8, np.full((i8.shape[0], 1), -1, np.int8)], axis=1)
                    vals = np.maximum(i8.astype(f32) / f32(127), f32(-1))
                    factor = f32(1.0 / 127)
                    lo = [f32(i8[:, c].min()) * factor if c < ch else f32(127) * factor for c in range(3)]
                    hi = [f32(i8[:, c].max()) * factor if c < ch else f32(-128) * factor for c in range(3)]
                    minmax = (np.array(lo, f32), np.array(hi, f32))
                elif comp == "float":
                    vals = np.array(tok[5:5 + w * h * d * ch], dtype=np.float64).astype(f32).reshape(d * h * w, ch)
                    if ch == 3:
                        vals = np.concatenate([vals, np.ones((vals.shape[0], 1), f32)], axis=1)
                else:
                    raise ValueError(f"unknown component type: {comp}")
                vals = np.ascontiguousarray(vals, f32)
                img = self.L.orc_add_image(self.h, w, h, d, 1 if ch == 1 else 4, _ptr(vals, c_float), 0, au, av, aw, linear, _ptr(border, c_float))
                if minmax is not None:
                    self.L.orc_set_image_minmax(self.h, img, _ptr(minmax[0], c_float), _ptr(minmax[1], c_float))
                    self.L.orc_set_image_border_levels(self.h, img, 127)   # SNorm8 storage
                t = self.L.orc_add_texture3d_array(self.h, img)
            elif ty in ("add", "mul"):  # texture3d/binary.cpp:150-160 (lhs is created first)
                lhs = self.add_texture3d(node.child("lhs"))
                rhs = self.add_texture3d(node.child("rhs"))
                t = self.L.orc_add_texture3d_binary(self.h, 0 if ty == "add" else 1, lhs, rhs)
            else:
                raise ValueError(f"unknown texture3d type: {ty}")
        self._pool[key] = t
        return t

    def add_medium(self, node: Node, world_to_local=None):
        """Returns a handle (index into self.media); ids are assigned at commit in priority order (core/scene.cpp:136-161)."""
        key = ("medium", id(node.raw))
        if key in self._pool and world_to_local is None:
            return self._pool[key]
        ty = node.child("type").as_str()
        if ty == "transform":  # factory/context.h:146-165, core/medium.cpp:98-129
            if node.find("world_to_local") is not None:
                w2l = node.child("world_to_local").as_transform3d().mat
            else:
                w2l = node.child("local_to_world").as_transform3d().inv
            if world_to_local is not None:  # nested: inner(outer(p))
                w2l = _matmul(w2l, world_to_local)
            h = self.add_medium(node.child("medium"), w2l)
        elif ty == "homogeneous":  # homogeneous.cpp:80-92
            prio = float(node.float_or("priority", 0.0))
            sigma_t, albedo, g = node.child("sigma_t").as_float(), node.child("albedo").as_vec3(), node.float_or("g", 0.0)
            self.media.append((prio, lambda: self.L.orc_add_medium_homogeneous(self.h, c_float(float(sigma_t)), _ptr(np.ascontiguousarray(albedo, f32), c_float), c_float(float(g)))))
            h = len(self.media) - 1
        elif ty == "hetergeneous":  # hetergeneous.cpp:113-135
            prio = float(node.float_or("priority", 0.0))
            st = self.add_texture3d(node.child("sigma_t"))
            al = self.add_texture3d(node.child("albedo"))
            g = self.add_texture3d(node.child("g")) if node.find("g") is not None else self.L.orc_add_texture3d_constant(self.h, 0, 0, 0)
            w = None if world_to_local is None else np.ascontiguousarray(world_to_local, f32)
            self.media.append((prio, lambda: self.L.orc_add_medium_heterogeneous(self.h, st, al, g, _ptr(w, c_float))))
            h = len(self.media) - 1
        else:
            raise ValueError(f"unknown medium type: {ty}")
        if world_to_local is None:
            self._pool[key] = h
        return h

    def add_instance(self, mesh, material, xf: Transform3D | None = None, intensity=None, inner_medium=None, outer_medium=None):
        """inner_medium / outer_medium: handles from add_medium (resolved to ids at commit) or None."""
        xf = xf or Transform3D()
        mat = np.ascontiguousarray(xf.mat, f32)
        inv = np.ascontiguousarray(xf.inv, f32)
        it = None if intensity is None else np.ascontiguousarray(intensity, f32)
        self.instances.append(dict(mesh=mesh, material=material, mat=mat, inv=inv, intensity=it, inner=inner_medium, outer=outer_medium))

    def set_gradient_sky(self, lower, upper, up):
        lo, hi, u = (np.ascontiguousarray(v, f32) for v in (lower, upper, up))
        self.L.orc_set_gradient_sky(self.h, _ptr(lo, c_float), _ptr(hi, c_float), _ptr(u, c_float))

    def set_camera(self, eye, dst, up, fov_y_deg):
        e, d, u = (np.ascontiguousarray(v, f32) for v in (eye, dst, up))
        self.L.orc_set_camera(self.h, _ptr(e, c_float), _ptr(d, c_float), _ptr(u, c_float), c_float(float(fov_y_deg)))

    def set_film(self, w, h):
        self.width, self.height = w, h
        self.L.orc_set_film(self.h, w, h)

    def set_renderer(self, spp=128, min_depth=5, max_depth=10, rr_threshold=0.1, rr_cont_prob=0.5, tile=False,
                     albedo=False, normal=False, filter_type=0, radius=1.3, alpha=1.6):
        self.spp = spp
        self.want_albedo, self.want_normal = albedo, normal
        self.L.orc_set_renderer(self.h, spp, min_depth, max_depth, c_float(float(rr_threshold)), c_float(float(rr_cont_prob)),
                                int(tile), int(albedo), int(normal), filter_type, c_float(float(radius)), c_float(float(alpha)))

    def commit(self):
        # medium ids: descending priority (stable for ties: creation order), Scene::commit (core/scene.cpp:136-161)
        order = sorted(range(len(self.media)), key=lambda i: -self.media[i][0])
        ids = {}
        for h in order:
            ids[h] = self.media[h][1]()
        for inst in self.instances:
            it = inst["intensity"]
            self.L.orc_add_instance(self.h, inst["mesh"], inst["material"], _ptr(inst["mat"], c_float), _ptr(inst["inv"], c_float),
                                    0 if it is None else 1, _ptr(it, c_float),
                                    -1 if inst["inner"] is None else ids[inst["inner"]], -1 if inst["outer"] is None else ids[inst["outer"]])
        self.L.orc_commit(self.h)

    def _load(self, filename):
        root = parse_scene_file(filename)
        self.scene_dir = os.path.dirname(os.path.abspath(filename))
        pool = self._pool
        scene = root.child("scene")
        for ent in scene.child("entities").elements():
            g = ent.child("geometry")
            if g.child("type").as_str() != "triangle_mesh":
                raise ValueError("unknown geometry type: " + g.child("type").as_str())
            if ("geo", id(g.raw)) not in pool:
                pos, nor, has_nor, uv = load_obj(self.resolve(g.child("filename").as_str()))
                read_pos = None
                if g.bool_or("transform_to_unit_cube", False):
                    read_pos, pos = pos, _to_unit_cube(pos, _OBJ_VERTICES[self.resolve(g.child("filename").as_str())])
                pool[("geo", id(g.raw))] = self.add_mesh_arrays(pos, nor, has_nor, uv, read_pos)
            m = ent.child("material")
            if ("mat", id(m.raw)) not in pool:
                pool[("mat", id(m.raw))] = self.add_material(m)
            xf = ent.child("local_to_world").as_transform3d() if ent.find("local_to_world") is not None else Transform3D()
            inten = None
            ln = ent.find("light")
            if ln is not None:
                if ln.child("type").as_str() != "mesh_light":
                    raise ValueError("area light expected")
                inten = ln.child("intensity").as_vec3()
            inner = self.add_medium(ent.child("inner_medium")) if ent.find("inner_medium") is not None else None
            outer = self.add_medium(ent.child("outer_medium")) if ent.find("outer_medium") is not None else None
            self.add_instance(pool[("geo", id(g.raw))], pool[("mat", id(m.raw))], xf, inten, inner, outer)
        if scene.find("volumes") is not None:  # factory/scene.cpp:57-80
            vols = scene.child("volumes")
            if not vols.is_array():
                raise ValueError("'volumes' is not an array node")
            for vol in vols.elements():
                geo = [np.ascontiguousarray(vol.child(k).as_vec3(), f32) for k in ("o", "x", "y", "z")]
                st = self.add_texture3d(vol.child("sigma_t"))
                al = self.add_texture3d(vol.child("albedo"))
                self.L.orc_add_volume(self.h, *[_ptr(g, c_float) for g in geo], st, al)
        env = scene.find("envir_light")
        if env is not None:
            ty = env.child("type").as_str()
            if ty == "gradient_sky":  # gradient_sky.cpp:56-67
                self.set_gradient_sky(env.child("lower").as_vec3(), env.child("upper").as_vec3(), env.child("up").as_vec3())
            elif ty == "ibl":  # ibl.cpp:89-102
                tex = self.add_texture(env.child("tex"))
                up = np.ascontiguousarray(env.child("up").as_vec3() if env.find("up") is not None else [0, 0, 1], f32)
                self.lut = (env.int_or("lut_width", 400), env.int_or("lut_height", 200))
                self.L.orc_set_ibl(self.h, tex, _ptr(up, c_float), self.lut[0], self.lut[1])
            else:
                raise ValueError(f"unknown light type: {ty}")
        w, h = root.child("width").as_int(), root.child("height").as_int()
        self.set_film(w, h)
        cam = root.child("camera")
        self.set_camera(cam.child("eye").as_vec3(), cam.child("dst").as_vec3(), cam.child("up").as_vec3(), cam.child("fov_y").as_degree())
        r = root.child("renderer")
        ftype, radius, alpha = 0, 1.3, 1.6
        fn = r.find("filter")
        if fn is not None and fn.child("type").as_str() == "gaussian":
            ftype, radius, alpha = 1, fn.float_or("radius", 1.3), fn.float_or("alpha", 1.6)
        self.set_renderer(spp=r.int_or("spp", 128), min_depth=r.int_or("min_depth", 5), max_depth=r.int_or("max_depth", 10),
                          rr_threshold=r.float_or("rr_threshold", 0.1), rr_cont_prob=r.float_or("rr_cont_prob", 0.5),
                          tile=r.bool_or("tile", False), albedo=r.bool_or("albedo", False), normal=r.bool_or("normal", False),
                          filter_type=ftype, radius=radius, alpha=alpha)
        sn = r.find("sampler")
        if sn is not None:
            if sn.as_str() not in ("independent", "halton"):
                raise ValueError("unknown sampler type: " + sn.as_str())
            self.L.orc_set_sampler(self.h, 1 if sn.as_str() == "halton" else 0)
        self.commit()

    # -- queries
    def num_triangles(self): return self.L.orc_num_triangles(self.h)

    def triangles(self):
        n = self.num_triangles()
        verts = np.zeros((n, 9), f32)
        ip = np.zeros((n, 2), np.uint32)
        self.L.orc_get_triangles(self.h, _ptr(verts, c_float), _ptr(ip, c_uint32))
        return verts, ip

    def mesh_tables(self, mesh_id):
        ntri = self.mesh_data[mesh_id][0].shape[0] // 3
        out = np.zeros((ntri, 24), f32)
        self.L.orc_get_mesh_tables(self.h, mesh_id, _ptr(out, c_float))
        return out

    def closest_hit(self, o_med, d_t1, brute=False):
        o_med = np.ascontiguousarray(o_med, f32); d_t1 = np.ascontiguousarray(d_t1, f32)
        n = o_med.shape[0]
        flag = np.zeros(n, np.uint32); detail = np.zeros((n, 4), np.uint32)
        self.L.orc_closest_hit(self.h, n, _ptr(o_med, c_float), _ptr(d_t1, c_float), _ptr(flag, c_uint32), _ptr(detail, c_uint32), 1 if brute else 0)
        return flag, detail

    def any_hit(self, o_med, d_t1):
        o_med = np.ascontiguousarray(o_med, f32); d_t1 = np.ascontiguousarray(d_t1, f32)
        n = o_med.shape[0]
        out = np.zeros(n, np.uint8)
        self.L.orc_any_hit(self.h, n, _ptr(o_med, c_float), _ptr(d_t1, c_float), _ptr(out, c_uint8))
        return out

    def generate(self, first, n):
        st = new_state(n)
        self.L.orc_generate(self.h, first, n, *state_ptrs(st))
        return st

    def shade(self, st, flag, detail):
        n = st["o_med"].shape[0]
        nxt = new_state(n)
        out = dict(status=np.zeros(n, np.uint32), next=nxt,
                   sh_o_med=np.zeros((n, 4), f32), sh_d_t1=np.zeros((n, 4), f32), sh_beta_li=np.zeros((n, 4), f32),
                   splat=np.zeros((n, 3), f32), albedo=np.zeros((n, 3), f32), normal=np.zeros((n, 3), f32))
        flag = np.ascontiguousarray(flag, np.uint32); detail = np.ascontiguousarray(detail, np.uint32)
        self.L.orc_shade(self.h, n, *state_ptrs(st), _ptr(flag, c_uint32), _ptr(detail, c_uint32), _ptr(out["status"], c_uint32),
                         *state_ptrs(nxt), _ptr(out["sh_o_med"], c_float), _ptr(out["sh_d_t1"], c_float), _ptr(out["sh_beta_li"], c_float),
                         _ptr(out["splat"], c_float), _ptr(out["albedo"], c_float), _ptr(out["normal"], c_float))
        return out

    def medium(self, st, flag, detail):
        """Medium stage (wavefront/medium.cpp).  Returns (updated state dict, updated flag, out dict like shade())."""
        n = st["o_med"].shape[0]
        st = {k: np.ascontiguousarray(v.copy()) for k, v in st.items()}
        flag = np.ascontiguousarray(flag.copy(), np.uint32); detail = np.ascontiguousarray(detail, np.uint32)
        nxt = new_state(n)
        out = dict(status=np.zeros(n, np.uint32), next=nxt, sh_o_med=np.zeros((n, 4), f32), sh_d_t1=np.zeros((n, 4), f32),
                   sh_beta_li=np.zeros((n, 4), f32), splat=np.zeros((n, 3), f32))
        self.L.orc_medium(self.h, n, *state_ptrs(st), _ptr(flag, c_uint32), _ptr(detail, c_uint32), _ptr(out["status"], c_uint32),
                          *state_ptrs(nxt), _ptr(out["sh_o_med"], c_float), _ptr(out["sh_d_t1"], c_float), _ptr(out["sh_beta_li"], c_float),
                          _ptr(out["splat"], c_float))
        return st, flag, out

    def shadow_tr(self, o_med, d_t1):
        o_med = np.ascontiguousarray(o_med, f32); d_t1 = np.ascontiguousarray(d_t1, f32)
        n = o_med.shape[0]
        tr = np.zeros((n, 3), f32)
        self.L.orc_shadow_tr(self.h, n, _ptr(o_med, c_float), _ptr(d_t1, c_float), _ptr(tr, c_float))
        return tr

    def has_medium(self): return bool(self.L.orc_has_medium(self.h))

    def vol_prim_medium_id(self): return int(self.L.orc_vol_prim_medium_id(self.h))

    def ibl_table(self):
        t = np.zeros((self.lut[1], self.lut[0]), f32)
        self.L.orc_get_ibl_table(self.h, _ptr(t, c_float))
        return t

    def eval_texture2d(self, tex, uv):
        uv = np.ascontiguousarray(uv, f32)
        out = np.zeros((uv.shape[0], 3), f32)
        self.L.orc_eval_texture2d(self.h, tex, uv.shape[0], _ptr(uv, c_float), _ptr(out, c_float))
        return out

    def bsdf_probe(self, material, interp_z, uv, wo, wi, sam):
        wo, wi, sam = (np.ascontiguousarray(a, f32) for a in (wo, wi, sam))
        n = wo.shape[0]
        iz = np.ascontiguousarray(interp_z, f32); uvv = np.ascontiguousarray(uv, f32)
        out = dict(sample=np.zeros((n, 8), f32), eval=np.zeros((n, 3), f32), pdf=np.zeros(n, f32), albedo=np.zeros(3, f32), normal=np.zeros(3, f32))
        self.L.orc_bsdf_probe(self.h, material, _ptr(iz, c_float), _ptr(uvv, c_float), n, _ptr(wo, c_float), _ptr(wi, c_float), _ptr(sam, c_float),
                              _ptr(out["sample"], c_float), _ptr(out["eval"], c_float), _ptr(out["pdf"], c_float), _ptr(out["albedo"], c_float),
                              _ptr(out["normal"], c_float))
        return out

    def render(self, spp_begin=0, spp_end=None, raw=False):
        spp_end = self.spp if spp_end is None else spp_end
        n = self.width * self.height
        color = np.zeros((self.height, self.width, 4), f32)
        albedo = np.zeros((self.height, self.width, 4), f32) if self.want_albedo else None
        normal = np.zeros((self.height, self.width, 4), f32) if self.want_normal else None
        rr = np.zeros((self.height, self.width, 4), f32) if raw else None
        rw = np.zeros((self.height, self.width), f32) if raw else None
        stats = np.zeros(7, np.uint64)
        self.L.orc_render(self.h, spp_begin, spp_end, _ptr(color, c_float), _ptr(albedo, c_float), _ptr(normal, c_float),
                          _ptr(rr, c_float), _ptr(rw, c_float), _ptr(stats, c_uint64))
        names = ["n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"]
        return dict(color=color, albedo=albedo, normal=normal, raw_radiance=rr, raw_weight=rw,
                    stats={k: int(v) for k, v in zip(names, stats)})


def _render_pixels(self, pixel_index, spp_begin=0, spp_end=None):
    """Raw film (radiance sum rgb, weight) of the listed pixels (index = y * width + x) over sample indices [spp_begin, spp_end)."""
    spp_end = self.spp if spp_end is None else spp_end
    idx = np.ascontiguousarray(pixel_index, np.uint32)
    out = np.zeros((idx.shape[0], 4), f32)
    stats = np.zeros(7, np.uint64)
    self.L.orc_render_pixels(self.h, c_int(idx.shape[0]), _ptr(idx, c_uint32), c_int(spp_begin), c_int(spp_end), _ptr(out, c_float), _ptr(stats, c_uint64))
    return out, {k: int(v) for k, v in zip(["n_gen", "n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss", "n_splat"], stats)}


OracleScene.render_pixels = _render_pixels


def new_state(n):
    return dict(o_med=np.zeros((n, 4), f32), d_t1=np.zeros((n, 4), f32), pixel=np.zeros((n, 2), np.uint32),
                beta_depth=np.zeros((n, 4), f32), path_rad=np.zeros((n, 4), f32), rng=np.zeros((n, 2), np.uint64),
                ble=np.zeros((n, 4), f32))


def state_ptrs(st):
    return [_ptr(st["o_med"], c_float), _ptr(st["d_t1"], c_float), _ptr(st["pixel"], c_uint32), _ptr(st["beta_depth"], c_float),
            _ptr(st["path_rad"], c_float), _ptr(st["rng"], c_uint64), _ptr(st["ble"], c_float)]


def _to_unit_cube(pos, vertices):
    """triangle_mesh_loader.cpp:193-213.  The loader still holds the file's indexed vertex list when this runs, so the bounding box is that
    of EVERY "v" line, referenced by a face or not (pinned against the reference's loader: tests/test_reference_host_pins_cpu.py, "random_unit_cube")."""
    lo, hi = vertices.min(axis=0), vertices.max(axis=0)
    ext = hi - lo
    m = ext.max()
    scale = f32(1) / m if m > 0 else f32(1)
    return (scale * (pos - f32(0.5) * (lo + hi))).astype(f32)


def load_obj_fan(filename):
    """lib/bvh/test/obj.hpp:57-95: positions only, negative indices, fan triangulation. Returns (ntri*3, 3)."""
    vs, tris = [], []
    with open(filename) as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "v":
                vs.append([f32(float(t)) for t in tok[1:4]])
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    i = int(t.split("/")[0])
                    idx.append(len(vs) + i if i < 0 else i - 1)
                for k in range(2, len(idx)):
                    tris += [vs[idx[0]], vs[idx[k - 1]], vs[idx[k]]]
    return np.array(tris, f32)
