"""Synthetic scene generators for the BASELINE.json configurations.

The reference ships no scene files, so every benchmark / parity scene is synthesised
here, deterministically, in Btrc's own JSON + OBJ formats (schema: SURVEY Appendix A).  The files
written are plain Btrc scene files: the reference's `BtrcCLI` would load them unchanged.
"""
from __future__ import annotations

import json
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORNELL_OBJ = os.path.join(_ROOT, "scenes", "assets", "cornell_box.obj")


def _fmt(x) -> str:
    return "%.9g" % float(x)


def write_obj(path, verts, faces, uvs=None, normals=None):
    """faces: list of index tuples (0-based) into verts; uvs/normals share the vertex indexing."""
    with open(path, "w") as f:
        for v in verts:
            f.write("v %s %s %s\n" % (_fmt(v[0]), _fmt(v[1]), _fmt(v[2])))
        if uvs is not None:
            for t in uvs:
                f.write("vt %s %s\n" % (_fmt(t[0]), _fmt(t[1])))
        if normals is not None:
            for n in normals:
                f.write("vn %s %s %s\n" % (_fmt(n[0]), _fmt(n[1]), _fmt(n[2])))
        for face in faces:
            toks = []
            for i in face:
                k = i + 1
                if uvs is not None and normals is not None:
                    toks.append(f"{k}/{k}/{k}")
                elif uvs is not None:
                    toks.append(f"{k}/{k}")
                elif normals is not None:
                    toks.append(f"{k}//{k}")
                else:
                    toks.append(str(k))
            f.write("f " + " ".join(toks) + "\n")


def split_cornell_groups(obj_path=CORNELL_OBJ):
    """Split the Cornell box OBJ (scenes/assets/cornell_box.obj) into its eight objects.

    The file names each object with `usemtl` BEFORE its faces but places the `g` line AFTER them for the two boxes,
    so the split follows `usemtl`.  Returns an ordered dict name -> (verts, faces), faces re-indexed into the
    object's own vertex list.
    """
    verts = []
    groups = {}
    cur = None
    with open(obj_path) as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "v":
                verts.append(tuple(np.float32(float(t)) for t in tok[1:4]))
            elif tok[0] == "usemtl":
                cur = tok[1]
                groups.setdefault(cur, ([], []))
            elif tok[0] == "f":
                gv, gf = groups[cur]
                face = []
                for t in tok[1:]:
                    i = int(t.split("/")[0])
                    i = len(verts) + i if i < 0 else i - 1
                    gv.append(verts[i])
                    face.append(len(gv) - 1)
                gf.append(tuple(face))
    return groups


def make_cornell(out_dir, width=512, height=512, spp=16, state_count=1000000, min_depth=5, max_depth=10,
                 filter=None, albedo=False, normal=False, tile=False, save=None):
    """BASELINE config 1: Cornell box, diffuse + one area light.  Returns the scene json path."""
    os.makedirs(out_dir, exist_ok=True)
    groups = split_cornell_groups()
    white, red, green = [0.725, 0.71, 0.68], [0.63, 0.065, 0.05], [0.14, 0.45, 0.091]
    colors = {"floor": white, "ceiling": white, "backWall": white, "rightWall": green, "leftWall": red,
              "shortBox": white, "tallBox": white, "light": [0.78, 0.78, 0.78]}
    entities = []
    for name, (gv, gf) in groups.items():
        write_obj(os.path.join(out_dir, f"{name}.obj"), gv, gf)
        ent = {
            "geometry": {"type": "triangle_mesh", "filename": "$env{scene_directory}/%s.obj" % name},
            "material": {"type": "diffuse", "color": colors[name]},
        }
        if name == "light":
            ent["light"] = {"type": "mesh_light", "intensity": [17, 12, 4]}
        entities.append(ent)
    renderer = {"type": "wfpt", "spp": spp, "min_depth": min_depth, "max_depth": max_depth, "rr_threshold": 0.1,
                "rr_cont_prob": 0.5, "state_count": state_count, "tile": tile, "albedo": albedo, "normal": normal,
                "filter": filter or {"type": "box"}}
    root = {
        "width": width, "height": height,
        "camera": {"type": "pinhole", "eye": [0, 1, 3.4], "dst": [0, 1, 0], "up": [0, 1, 0], "fov_y": {"deg": 39}},
        "renderer": renderer,
        "scene": {"entities": entities},
    }
    if save:
        root["post_processors"] = [{"type": "save", "color": save, "gamma": 2.2}]
    path = os.path.join(out_dir, "scene.json")
    with open(path, "w") as f:
        json.dump(root, f, indent=1)
    return path


# ---------------------------------------------------------------------------------------------------- mesh generators
def uv_sphere(rings, segments, radius=1.0, center=(0.0, 0.0, 0.0)):
    """UV sphere with per-vertex normals and texture coordinates: 2 * segments * (rings - 1) triangles, no degenerate ones."""
    verts, uvs, nors = [], [], []
    for r in range(rings + 1):
        theta = np.pi * r / rings
        for s in range(segments + 1):
            phi = 2 * np.pi * s / segments
            n = np.array([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)])
            verts.append(np.asarray(center) + radius * n)
            nors.append(n)
            uvs.append((s / segments, r / rings))
    faces = []
    row = segments + 1
    for r in range(rings):
        for s in range(segments):
            a, b, c, d = r * row + s, r * row + s + 1, (r + 1) * row + s + 1, (r + 1) * row + s
            if r != 0:
                faces.append((a, b, c))
            if r != rings - 1:
                faces.append((a, c, d))
    return verts, faces, uvs, nors


def quad(p0, p1, p2, p3):
    """Two triangles p0 p1 p2 / p0 p2 p3 with uv (0,0) (1,0) (1,1) (0,1)."""
    return [p0, p1, p2, p3], [(0, 1, 2), (0, 2, 3)], [(0, 0), (1, 0), (1, 1), (0, 1)]


def _smooth_noise(rng, res, cells):
    """Low-frequency value noise in [0, 1]: random grid, bilinearly upsampled (wraps)."""
    g = rng.random((cells, cells))
    x = np.arange(res) * cells / res
    i0 = np.floor(x).astype(int) % cells
    i1 = (i0 + 1) % cells
    f = x - np.floor(x)
    f = f * f * (3 - 2 * f)
    rows = g[i0][:, i0] * (1 - f)[None, :] + g[i0][:, i1] * f[None, :]
    rows1 = g[i1][:, i0] * (1 - f)[None, :] + g[i1][:, i1] * f[None, :]
    return rows * (1 - f)[:, None] + rows1 * f[:, None]


def write_test_textures(out_dir, res=1024, seed=7):
    """base colour (sRGB, RGB), roughness (single channel) and a tangent-space normal map, as 8-bit PNGs."""
    from PIL import Image

    rng = np.random.default_rng(seed)
    col = np.stack([0.15 + 0.8 * (0.6 * _smooth_noise(rng, res, 8) + 0.4 * _smooth_noise(rng, res, 32)) for _ in range(3)], axis=2)
    Image.fromarray((np.clip(col, 0, 1) * 255 + 0.5).astype(np.uint8), "RGB").save(os.path.join(out_dir, "base_color.png"))
    rough = 0.08 + 0.7 * _smooth_noise(rng, res, 16)
    Image.fromarray((np.clip(rough, 0, 1) * 255 + 0.5).astype(np.uint8), "L").save(os.path.join(out_dir, "roughness.png"))
    h = _smooth_noise(rng, res, 24)
    dx = np.roll(h, -1, axis=1) - np.roll(h, 1, axis=1)
    dy = np.roll(h, -1, axis=0) - np.roll(h, 1, axis=0)
    n = np.stack([-dx * res / 64, -dy * res / 64, np.ones_like(h)], axis=2)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    Image.fromarray((np.clip(0.5 + 0.5 * n, 0, 1) * 255 + 0.5).astype(np.uint8), "RGB").save(os.path.join(out_dir, "normal.png"))
    return "base_color.png", "roughness.png", "normal.png"


def _renderer(spp, state_count, min_depth, max_depth, filter, albedo=False, normal=False, tile=False):
    return {"type": "wfpt", "spp": spp, "min_depth": min_depth, "max_depth": max_depth, "rr_threshold": 0.1, "rr_cont_prob": 0.5,
            "state_count": state_count, "tile": tile, "albedo": albedo, "normal": normal, "filter": filter or {"type": "box"}}


def _geo(name):
    return {"type": "triangle_mesh", "filename": "$env{scene_directory}/%s.obj" % name}


def _dump(out_dir, root):
    path = os.path.join(out_dir, "scene.json")
    with open(path, "w") as f:
        json.dump(root, f, indent=1)
    return path


def make_material_test(out_dir, width=1920, height=1080, spp=256, state_count=1000000, rings=100, segments=100, tex_res=1024,
                       min_depth=5, max_depth=10, filter=None, grid=(5, 3), albedo=False, normal=False, save=None, sky=True):
    """BASELINE config 2: Disney BSDF material-test scene with image textures.

    Ground plane + grid of UV spheres (2*segments*(rings-1) triangles each; 19,800 at the defaults), every sphere a `disney`
    material sweeping metallic / roughness / clearcoat / sheen / transmission (+ one each of metal, glass, mirror so that every
    material type is on the path); base colour (sRGB) and roughness from generated PNG textures, one normal map; two quad
    mesh lights + gradient sky; gaussian film filter."""
    os.makedirs(out_dir, exist_ok=True)
    base, rough, nmap = write_test_textures(out_dir, tex_res)
    tex = lambda name, **kw: dict({"type": "array", "filename": "$env{scene_directory}/" + name, "address_mode_u": "wrap", "address_mode_v": "wrap"}, **kw)
    base_t = tex(base, srgb_to_linear=True)
    rough_t = tex(rough)
    v, f, uv, n = uv_sphere(rings, segments)
    write_obj(os.path.join(out_dir, "sphere.obj"), v, f, uv, n)
    gv, gf, guv = quad((-12, 0, 8), (12, 0, 8), (12, 0, -12), (-12, 0, -12))
    write_obj(os.path.join(out_dir, "ground.obj"), gv, gf, [(u * 6, w * 6) for u, w in guv])
    lv, lf, luv = quad((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1))  # geometric normal cross(b-a, c-a) = -y: emits downwards (mesh lights are one-sided)
    write_obj(os.path.join(out_dir, "lightquad.obj"), lv, lf, luv)
    entities = [{"geometry": _geo("ground"), "material": {"type": "disney", "color": base_t, "metallic": 0.0, "roughness": rough_t, "normal": tex(nmap)}}]
    nx, ny = grid
    sphere_geo = _geo("sphere")
    k = 0
    for j in range(ny):
        for i in range(nx):
            fx, fy = i / max(nx - 1, 1), j / max(ny - 1, 1)
            if j == 0:      # metallic sweep, textured roughness
                mat = {"type": "disney", "color": base_t, "metallic": fx, "roughness": rough_t, "specular_tint": 0.5 * fx}
            elif j == 1:    # roughness sweep with clearcoat / sheen / anisotropy
                mat = {"type": "disney", "color": [0.8, 0.25 + 0.5 * fx, 0.2], "metallic": 0.1, "roughness": 0.05 + 0.9 * fx,
                       "clearcoat": 1.0 - fx, "clearcoat_gloss": 0.9, "sheen": fx, "sheen_tint": 0.5, "anisotropic": 0.8 * fx}
            else:           # transmission sweep + the non-Disney materials
                if i == nx - 1:
                    mat = {"type": "glass", "color": [0.95, 0.97, 1.0], "ior": 1.5}
                elif i == nx - 2:
                    mat = {"type": "mirror", "color": [0.9, 0.85, 0.7]}
                elif i == nx - 3:
                    mat = {"type": "metal", "color": [0.95, 0.64, 0.54], "roughness": 0.25, "anisotropic": 0.5}
                else:
                    mat = {"type": "disney", "color": [0.9, 0.9, 0.95], "metallic": 0.0, "roughness": 0.1 + 0.3 * fx, "transmission": 0.9,
                           "ior": 1.45, "transmission_roughness": 0.05 + 0.3 * fx}
            x = (i - (nx - 1) / 2) * 2.4
            z = -(j * 2.6)
            entities.append({"geometry": sphere_geo, "material": mat,
                             "local_to_world": [{"rotate_y": {"deg": 30.0 * k}}, {"translate": [x, 1.0, z]}]})
            k += 1
    for lx, inten in ((-4.0, [40, 36, 30]), (4.0, [24, 30, 40])):
        entities.append({"geometry": _geo("lightquad"), "material": {"type": "diffuse", "color": [0.0, 0.0, 0.0]},
                         "local_to_world": [{"scale": [1.5, 1.0, 1.5]}, {"translate": [lx, 7.0, -2.0]}],
                         "light": {"type": "mesh_light", "intensity": inten}})
    scene = {"entities": entities}
    if sky:
        scene["envir_light"] = {"type": "gradient_sky", "upper": [0.35, 0.45, 0.7], "lower": [0.12, 0.1, 0.08], "up": [0, 1, 0]}
    root = {"width": width, "height": height,
            "camera": {"type": "pinhole", "eye": [0, 5.2, 9.5], "dst": [0, 0.6, -2.2], "up": [0, 1, 0], "fov_y": {"deg": 42}},
            "renderer": _renderer(spp, state_count, min_depth, max_depth, filter or {"type": "gaussian", "radius": 1.3, "alpha": 1.6}, albedo, normal),
            "scene": scene}
    if save:
        root["post_processors"] = [{"type": "tonemap", "exposure": 1.0}, {"type": "save", "color": save, "gamma": 2.2}]
    return _dump(out_dir, root)


def write_density_grid(path, res=128, seed=11, max_sigma=8.0):
    """Procedural cloud density as a Btrc text grid ("float 1 W H D", x fastest): fBm of value noise, soft spherical falloff."""
    rng = np.random.default_rng(seed)

    def noise3(cells):
        g = rng.random((cells, cells, cells))
        x = np.arange(res) * cells / res
        i0 = np.floor(x).astype(int) % cells
        i1 = (i0 + 1) % cells
        f = x - np.floor(x)
        f = f * f * (3 - 2 * f)
        out = 0
        for dz, wz in ((i0, 1 - f), (i1, f)):
            for dy, wy in ((i0, 1 - f), (i1, f)):
                for dx, wx in ((i0, 1 - f), (i1, f)):
                    out = out + g[dz][:, dy][:, :, dx] * wz[:, None, None] * wy[None, :, None] * wx[None, None, :]
        return out

    d = 0.55 * noise3(4) + 0.3 * noise3(8) + 0.15 * noise3(16)
    c = (np.arange(res) + 0.5) / res * 2 - 1
    r = np.sqrt(c[:, None, None] ** 2 + c[None, :, None] ** 2 + c[None, None, :] ** 2)
    d = np.clip((d - 0.35) * 2.2, 0, 1) * np.clip(1.15 - r, 0, 1)
    d = (d / d.max() * max_sigma).astype(np.float32)
    with open(path, "w") as f:
        f.write(f"float 1 {res} {res} {res}\n")
        np.savetxt(f, d.reshape(-1, res), fmt="%.5g")
    return d


def make_medium_scene(out_dir, width=1920, height=1080, spp=128, state_count=1000000, grid_res=128, min_depth=5, max_depth=10,
                      heterogeneous=True, filter=None, save=None):
    """BASELINE config 3: heterogeneous grid-volume medium (delta / ratio tracking) with NEE.

    The Cornell box plus a cube entity with the `invisible` material whose inner medium is transform -> hetergeneous
    (sigma_t from a generated text grid, albedo 0.9, g 0.3), or a homogeneous medium when heterogeneous=False."""
    os.makedirs(out_dir, exist_ok=True)
    groups = split_cornell_groups()
    white, red, green = [0.725, 0.71, 0.68], [0.63, 0.065, 0.05], [0.14, 0.45, 0.091]
    colors = {"floor": white, "ceiling": white, "backWall": white, "rightWall": green, "leftWall": red, "light": [0.78, 0.78, 0.78]}
    entities = []
    for name, (gv, gf) in groups.items():
        if name not in colors:
            continue  # the two boxes make room for the cloud
        write_obj(os.path.join(out_dir, f"{name}.obj"), gv, gf)
        ent = {"geometry": _geo(name), "material": {"type": "diffuse", "color": colors[name]}}
        if name == "light":
            ent["light"] = {"type": "mesh_light", "intensity": [17, 12, 4]}
        entities.append(ent)
    # unit cube [-1,1]^3 (12 triangles), outward normals
    c = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    faces = [(0, 2, 1), (0, 3, 2), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (3, 7, 6), (3, 6, 2), (0, 4, 7), (0, 7, 3), (1, 2, 6), (1, 6, 5)]
    write_obj(os.path.join(out_dir, "cube.obj"), c, faces)
    l2w = [{"scale": [0.55, 0.55, 0.55]}, {"rotate_y": {"deg": 20}}, {"translate": [0.0, 0.8, 0.0]}]
    if heterogeneous:
        write_density_grid(os.path.join(out_dir, "density.txt"), grid_res)
        medium = {"type": "transform", "local_to_world": l2w,
                  "medium": {"type": "hetergeneous", "sigma_t": {"type": "array", "text": "$env{scene_directory}/density.txt"},
                             "albedo": [0.9, 0.9, 0.9], "g": 0.3}}
    else:
        medium = {"type": "homogeneous", "sigma_t": 2.5, "albedo": [0.9, 0.85, 0.8], "g": 0.3}
    entities.append({"geometry": _geo("cube"), "material": {"type": "invisible"}, "local_to_world": l2w, "inner_medium": medium})
    root = {"width": width, "height": height,
            "camera": {"type": "pinhole", "eye": [0, 1, 3.4], "dst": [0, 1, 0], "up": [0, 1, 0], "fov_y": {"deg": 39}},
            "renderer": _renderer(spp, state_count, min_depth, max_depth, filter),
            "scene": {"entities": entities}}
    if save:
        root["post_processors"] = [{"type": "save", "color": save, "gamma": 2.2}]
    return _dump(out_dir, root)


def make_volume_scene(out_dir, width=640, height=480, spp=32, state_count=500000, grid_res=32, min_depth=5, max_depth=10, filter=None,
                      save=None, n_extra=0):
    """Volume primitives (the scene file's "volumes" array; core/volume.cpp): the open Cornell box with two overlapping
    parallelepipeds -- a rotated grid-density cloud and a constant-density slab that cuts through it -- plus `n_extra`
    small constant boxes (more than four of them overlap in one corner when n_extra >= 5)."""
    os.makedirs(out_dir, exist_ok=True)
    groups = split_cornell_groups()
    white, red, green = [0.725, 0.71, 0.68], [0.63, 0.065, 0.05], [0.14, 0.45, 0.091]
    colors = {"floor": white, "ceiling": white, "backWall": white, "rightWall": green, "leftWall": red, "light": [0.78, 0.78, 0.78]}
    entities = []
    for name, (gv, gf) in groups.items():
        if name not in colors:
            continue
        write_obj(os.path.join(out_dir, f"{name}.obj"), gv, gf)
        ent = {"geometry": _geo(name), "material": {"type": "diffuse", "color": colors[name]}}
        if name == "light":
            ent["light"] = {"type": "mesh_light", "intensity": [17, 12, 4]}
        entities.append(ent)
    write_density_grid(os.path.join(out_dir, "density.txt"), grid_res, max_sigma=6.0)
    c, s_ = float(np.cos(np.radians(25.0))), float(np.sin(np.radians(25.0)))
    volumes = [
        {"o": [-0.55, 0.35, -0.35], "x": [0.9 * c, 0.0, -0.9 * s_], "y": [0.0, 0.9, 0.0], "z": [0.9 * s_, 0.0, 0.9 * c],
         "sigma_t": {"type": "array", "text": "$env{scene_directory}/density.txt"}, "albedo": [0.9, 0.9, 0.95]},
        {"o": [-0.9, 0.6, -0.6], "x": [1.8, 0.0, 0.0], "y": [0.0, 0.25, 0.0], "z": [0.0, 0.0, 1.2],
         "sigma_t": 1.5, "albedo": [0.95, 0.6, 0.3]},
    ]
    for k in range(n_extra):
        volumes.append({"o": [0.35 + 0.02 * k, 1.2 + 0.015 * k, -0.2 + 0.01 * k], "x": [0.4, 0, 0], "y": [0, 0.4, 0], "z": [0, 0, 0.4],
                        "sigma_t": 0.8 + 0.3 * k, "albedo": [0.5 + 0.05 * k, 0.8, 0.9 - 0.05 * k]})
    root = {"width": width, "height": height,
            "camera": {"type": "pinhole", "eye": [0, 1, 3.4], "dst": [0, 1, 0], "up": [0, 1, 0], "fov_y": {"deg": 39}},
            "renderer": _renderer(spp, state_count, min_depth, max_depth, filter),
            "scene": {"entities": entities, "volumes": volumes}}
    if save:
        root["post_processors"] = [{"type": "save", "color": save, "gamma": 2.2}]
    return _dump(out_dir, root)


def write_envmap(path, width=512, height=256, seed=5):
    """Equirectangular environment map as an 8-bit PNG: sky gradient, a bright sun blob and soft clouds (u = phi / 2pi, v = theta / pi)."""
    from PIL import Image

    rng = np.random.default_rng(seed)
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    sky = np.stack([0.25 + 0.3 * (1 - v), 0.35 + 0.35 * (1 - v), 0.55 + 0.4 * (1 - v)], axis=1)[:, None, :] * np.ones((1, width, 1))
    ground = np.array([0.18, 0.15, 0.12])
    img = np.where((v > 0.5)[:, None, None], ground[None, None, :] * np.ones((height, width, 1)), sky)
    clouds = _smooth_noise(rng, max(width, height), 8)[:height, :width]
    img[: height // 2] += 0.25 * clouds[: height // 2, :, None]
    du = np.minimum(np.abs(u - 0.3), 1 - np.abs(u - 0.3))[None, :]
    dv = (v - 0.22)[:, None]
    sun = np.exp(-((du * 2) ** 2 + dv ** 2) / (2 * 0.02 ** 2))
    img += sun[:, :, None] * np.array([4.0, 3.6, 3.0])
    Image.fromarray((np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8), "RGB").save(path)


def make_ibl_scene(out_dir, width=640, height=360, spp=64, state_count=500000, rings=24, segments=32, env_res=(512, 256), lut=(64, 32),
                   min_depth=5, max_depth=10, save=None):
    """Image-based lighting (light/ibl.cpp): ground + three spheres lit by an environment map only; the map is an 8-bit PNG
    scaled by a `mul` operator texture (texture2d/math.cpp) so that the sun is brighter than 1."""
    os.makedirs(out_dir, exist_ok=True)
    write_envmap(os.path.join(out_dir, "env.png"), *env_res)
    v, f, uv, n = uv_sphere(rings, segments)
    write_obj(os.path.join(out_dir, "sphere.obj"), v, f, uv, n)
    gv, gf, guv = quad((-8, 0, 8), (8, 0, 8), (8, 0, -8), (-8, 0, -8))
    write_obj(os.path.join(out_dir, "ground.obj"), gv, gf, guv)
    mats = [{"type": "diffuse", "color": [0.7, 0.7, 0.7]},
            {"type": "metal", "color": [0.95, 0.8, 0.5], "roughness": 0.2},
            {"type": "disney", "color": [0.2, 0.5, 0.8], "metallic": 0.0, "roughness": 0.4, "clearcoat": 0.5}]
    entities = [{"geometry": _geo("ground"), "material": {"type": "diffuse", "color": [0.5, 0.5, 0.5]}}]
    for i, m in enumerate(mats):
        entities.append({"geometry": _geo("sphere"), "material": m, "local_to_world": {"translate": [(i - 1) * 2.4, 1.0, 0.0]}})
    env_tex = {"type": "mul", "lhs": {"type": "array", "filename": "$env{scene_directory}/env.png", "address_mode_u": "wrap", "srgb_to_linear": True},
               "rhs": [6.0, 6.0, 6.0]}
    root = {"width": width, "height": height,
            "camera": {"type": "pinhole", "eye": [0, 2.5, 7.5], "dst": [0, 0.8, 0], "up": [0, 1, 0], "fov_y": {"deg": 40}},
            "renderer": _renderer(spp, state_count, min_depth, max_depth, None),
            "scene": {"entities": entities,
                      "envir_light": {"type": "ibl", "tex": env_tex, "up": [0, 1, 0], "lut_width": lut[0], "lut_height": lut[1]}}}
    if save:
        root["post_processors"] = [{"type": "tonemap", "exposure": 1.0}, {"type": "save", "color": save, "gamma": 2.2}]
    return _dump(out_dir, root)


def blob_mesh(rings, segments, seed=13, amplitude=0.18):
    """A bumpy sphere: the UV sphere displaced along its normals by low-frequency value noise (2*segments*(rings-1) triangles).
    Per-vertex normals are those of the underlying sphere (smooth shading across the bumps, as with a normal-less scan)."""
    rng = np.random.default_rng(seed)
    verts, faces, uvs, nors = uv_sphere(rings, segments)
    noise = _smooth_noise(rng, max(rings, segments) + 1, 6)
    out = []
    row = segments + 1
    for i, (v, n) in enumerate(zip(verts, nors)):
        r, s = divmod(i, row)
        s = s % segments  # the seam column repeats column 0
        pole = np.sin(np.pi * r / rings)  # no displacement at the poles (all seam vertices coincide there)
        out.append(np.asarray(n) * (1.0 + amplitude * pole * (noise[r % noise.shape[0], s % noise.shape[1]] - 0.5) * 2.0))
    return out, faces, uvs, nors


def make_instanced_scene(out_dir, width=1920, height=1080, spp=128, state_count=1000000, rings=101, segments=200, n_instances=250,
                         n_lights=64, min_depth=5, max_depth=16, seed=13, filter=None, save=None):
    """BASELINE config 4: one mesh (2*segments*(rings-1) = 40,000 triangles at the defaults) instanced n_instances times with
    random rigid transforms (+ uniform scale) inside a closed room, mixed diffuse / metal materials, n_lights quad mesh-light
    entities under the ceiling, deep paths (max_depth 16) (250 x 40,000 = 10 M triangles, 64 lights).

    The mesh is written once; every entity after the first refers to the first one's geometry node with `$reference{...}`, so the
    factory pools it (factory/context.h:205-236) and the OBJ is loaded once."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    v, f, uv, n = blob_mesh(rings, segments, seed)
    write_obj(os.path.join(out_dir, "blob.obj"), v, f, uv, n)
    X, Y, Z = 20.0, 9.0, 20.0  # room half-extent x, height, half-extent z
    room = {
        "floor": quad((-X, 0, Z), (X, 0, Z), (X, 0, -Z), (-X, 0, -Z)),
        "ceil": quad((-X, Y, -Z), (X, Y, -Z), (X, Y, Z), (-X, Y, Z)),
        "back": quad((-X, 0, -Z), (X, 0, -Z), (X, Y, -Z), (-X, Y, -Z)),
        "front": quad((X, 0, Z), (-X, 0, Z), (-X, Y, Z), (X, Y, Z)),
        "left": quad((-X, 0, Z), (-X, 0, -Z), (-X, Y, -Z), (-X, Y, Z)),
        "right": quad((X, 0, -Z), (X, 0, Z), (X, Y, Z), (X, Y, -Z)),
    }
    wall_col = {"floor": [0.6, 0.6, 0.6], "ceil": [0.75, 0.75, 0.75], "back": [0.7, 0.7, 0.65], "front": [0.7, 0.7, 0.65],
                "left": [0.65, 0.2, 0.15], "right": [0.2, 0.5, 0.25]}
    entities = []
    for name, (qv, qf, quv) in room.items():
        write_obj(os.path.join(out_dir, f"{name}.obj"), qv, qf, quv)
        entities.append({"geometry": _geo(name), "material": {"type": "diffuse", "color": wall_col[name]}})
    first_blob = len(entities)
    for k in range(n_instances):
        scale = float(rng.uniform(0.7, 1.6))
        pos = [float(rng.uniform(-X + 2, X - 2)), float(rng.uniform(1.0, Y - 2.5)), float(rng.uniform(-Z + 2, Z - 6))]
        axis = rng.normal(size=3)
        axis = (axis / np.linalg.norm(axis)).tolist()
        col = [float(c) for c in rng.uniform(0.25, 0.9, 3)]
        if k % 3 == 2:
            mat = {"type": "metal", "color": col, "roughness": float(rng.uniform(0.1, 0.5))}
        else:
            mat = {"type": "diffuse", "color": col}
        geo = _geo("blob") if k == 0 else "$reference{$root/scene/entities/%d/geometry}" % first_blob
        entities.append({"geometry": geo, "material": mat,
                         "local_to_world": [{"scale": [scale, scale, scale]}, {"rotate": {"axis": axis, "angle": {"deg": float(rng.uniform(0, 360))}}},
                                            {"translate": pos}]})
    lv, lf, luv = quad((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1))  # geometric normal cross(b-a, c-a) = -y: emits downwards (mesh lights are one-sided)
    write_obj(os.path.join(out_dir, "lightquad.obj"), lv, lf, luv)
    side = int(np.ceil(np.sqrt(n_lights)))
    first_light = len(entities)
    for k in range(n_lights):
        i, j = k % side, k // side
        lx = (i + 0.5) / side * 2 * (X - 1.5) - (X - 1.5)
        lz = (j + 0.5) / side * 2 * (Z - 1.5) - (Z - 1.5)
        inten = [float(c) for c in (14.0 + 8.0 * rng.random(3))]
        geo = _geo("lightquad") if k == 0 else "$reference{$root/scene/entities/%d/geometry}" % first_light
        entities.append({"geometry": geo, "material": {"type": "diffuse", "color": [0.0, 0.0, 0.0]},
                         "local_to_world": [{"scale": [0.8, 1.0, 0.8]}, {"translate": [lx, Y - 0.05, lz]}],
                         "light": {"type": "mesh_light", "intensity": inten}})
    root = {"width": width, "height": height,
            "camera": {"type": "pinhole", "eye": [0, 5.0, Z - 0.5], "dst": [0, 3.2, 0], "up": [0, 1, 0], "fov_y": {"deg": 55}},
            "renderer": _renderer(spp, state_count, min_depth, max_depth, filter),
            "scene": {"entities": entities}}
    if save:
        root["post_processors"] = [{"type": "tonemap", "exposure": 1.0}, {"type": "save", "color": save, "gamma": 2.2}]
    return _dump(out_dir, root)
