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
