"""Headless UV / angle / depth rendering of a UV-parametrised mesh: the reference's stand-alone OpenGL
renderer (`scripts/scannet/render_uv`, driven by `scripts/scannet/render_uvs.py`) on the CUDA rasteriser
(csrc/raster_kernels.cu, C-ABI `smb_raster_view`).

    mesh = load_mesh("scene0000_00_uvs_blender.ply")            # OBJ (v / vt / vn / f) or PLY (ascii / binary LE)
    r = MeshRasterizer(mesh)                                    # uploads once
    uv, angle, depth = r.render(pose_c2w, K, (Wk, Hk), (w, h))  # three (h, w, 3) float32 CUDA tensors
    render_scene(mesh_path, pose_dir, intrinsics_txt, scene_dir, ...)   # writes uv/, uv_<h>/ like render_uvs.py + main.cpp

    python -m stylemesh_b200.raster --mesh M --pose_dir P --intrinsics I --out SCENE_DIR [--multi_size ...]

Camera, shader and read-back semantics are those of the reference (file:line in oracle/raster_oracle.py and
csrc/raster_kernels.cu); like Assimp's aiProcess_FlipUVs (include/model.h:57) the loaders store v = 1 - v_file, and like
aiProcess_GenNormals they compute normals when the file has none.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi

NEAR, FAR, TEX_SIZE = 0.1, 10.0, 1024.0


@dataclass
class Mesh:
    verts: np.ndarray          # (V, 3) float32
    faces: np.ndarray          # (F, 3) int32
    corner_uv: np.ndarray      # (F, 3, 2) float32, v already flipped (1 - v_file)
    corner_normal: np.ndarray  # (F, 3, 3) float32


def _face_normals_to_corners(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """aiProcess_GenNormals: one (flat) normal per face, shared by its three corners."""
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    n = np.cross(b - a, c - a)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    return np.repeat(n[:, None, :], 3, axis=1).astype(np.float32)


def _finish(verts, faces, corner_uv, corner_normal, flip_v=True) -> Mesh:
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    if corner_uv is None:
        corner_uv = np.zeros((len(faces), 3, 2), np.float32)              # model.h:147 - no texture coordinates: (0, 0)
    corner_uv = np.array(corner_uv, dtype=np.float32)
    if flip_v:
        corner_uv[..., 1] = 1.0 - corner_uv[..., 1]
    if corner_normal is None:
        corner_normal = _face_normals_to_corners(verts, faces)
    return Mesh(verts, faces, np.ascontiguousarray(corner_uv), np.ascontiguousarray(corner_normal, dtype=np.float32))


def load_obj(path: str) -> Mesh:
    v, vt, vn, fv, ft, fn = [], [], [], [], [], []
    with open(path) as fh:
        for line in fh:
            s = line.split()
            if not s:
                continue
            if s[0] == "v":
                v.append([float(x) for x in s[1:4]])
            elif s[0] == "vt":
                vt.append([float(x) for x in s[1:3]])
            elif s[0] == "vn":
                vn.append([float(x) for x in s[1:4]])
            elif s[0] == "f":
                idx = [tok.split("/") for tok in s[1:]]
                for k in range(1, len(idx) - 1):                          # aiProcess_Triangulate: fan
                    tri = [idx[0], idx[k], idx[k + 1]]
                    fv.append([int(t[0]) for t in tri])
                    ft.append([int(t[1]) if len(t) > 1 and t[1] else 0 for t in tri])
                    fn.append([int(t[2]) if len(t) > 2 and t[2] else 0 for t in tri])
    v = np.asarray(v, np.float32)
    fix = lambda a, n: np.where(np.asarray(a) < 0, np.asarray(a) + n, np.asarray(a) - 1)      # 1-based / negative indices
    faces = fix(fv, len(v))
    cuv = np.asarray(vt, np.float32)[fix(ft, len(vt))] if vt and np.all(np.asarray(ft) != 0) else None
    cn = np.asarray(vn, np.float32)[fix(fn, len(vn))] if vn and np.all(np.asarray(fn) != 0) else None
    return _finish(v, faces, cuv, cn)


_PLY_TYPES = {"char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H",
              "uint16": "H", "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f",
              "double": "d", "float64": "d"}


def load_ply(path: str) -> Mesh:
    """Vertex properties x y z [nx ny nz] [s t | u v | texture_u texture_v]; face lists vertex_indices / vertex_index and
    optionally texcoord (six floats per triangle: per-corner UVs, as Blender / MeshLab write them)."""
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = fh.readline().decode("ascii", "replace").strip()
            if line == "end_header":
                break
            tok = line.split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
            elif tok[0] == "property":
                elements[-1]["props"].append(tok[1:])
        if fmt not in ("ascii", "binary_little_endian"):
            raise ValueError(f"{path}: PLY format {fmt} is not supported (ascii / binary_little_endian)")
        data = {}
        for el in elements:
            scalar = all(p[0] != "list" for p in el["props"])
            names = [p[-1] for p in el["props"]]
            if fmt == "binary_little_endian" and scalar:
                dt = np.dtype([(p[1], "<" + _PLY_TYPES[p[0]]) for p in el["props"]])
                arr = np.frombuffer(fh.read(dt.itemsize * el["count"]), dtype=dt)
                data[el["name"]] = {n: arr[n] for n in names}
                continue
            rows = {n: [] for n in names}
            for _ in range(el["count"]):
                if fmt == "ascii":
                    vals = fh.readline().split()
                    pos = 0
                for p in el["props"]:
                    if p[0] == "list":
                        if fmt == "ascii":
                            cnt = int(vals[pos]); pos += 1
                            item = [float(x) for x in vals[pos:pos + cnt]]; pos += cnt
                        else:
                            ct, it = _PLY_TYPES[p[1]], _PLY_TYPES[p[2]]
                            cnt = struct.unpack("<" + ct, fh.read(struct.calcsize(ct)))[0]
                            item = list(struct.unpack("<" + it * cnt, fh.read(struct.calcsize(it) * cnt)))
                        rows[p[-1]].append(item)
                    else:
                        if fmt == "ascii":
                            rows[p[-1]].append(float(vals[pos])); pos += 1
                        else:
                            t = _PLY_TYPES[p[0]]
                            rows[p[-1]].append(struct.unpack("<" + t, fh.read(struct.calcsize(t)))[0])
            data[el["name"]] = rows
    vx = data["vertex"]
    verts = np.stack([np.asarray(vx[k], np.float32) for k in ("x", "y", "z")], 1)
    fkey = "vertex_indices" if "vertex_indices" in data["face"] else "vertex_index"
    faces, corner_uv_list = [], []
    tex = data["face"].get("texcoord")
    for i, poly in enumerate(data["face"][fkey]):
        poly = [int(x) for x in poly]
        for k in range(1, len(poly) - 1):
            faces.append([poly[0], poly[k], poly[k + 1]])
            if tex is not None:
                t = np.asarray(tex[i], np.float32).reshape(-1, 2)
                corner_uv_list.append([t[0], t[k], t[k + 1]])
    faces = np.asarray(faces, np.int64)
    cuv = None
    if tex is not None:
        cuv = np.asarray(corner_uv_list, np.float32)
    else:
        for a, b in (("s", "t"), ("u", "v"), ("texture_u", "texture_v")):
            if a in vx and b in vx:
                cuv = np.stack([np.asarray(vx[a], np.float32), np.asarray(vx[b], np.float32)], 1)[faces]
                break
    cn = None
    if all(k in vx for k in ("nx", "ny", "nz")):
        cn = np.stack([np.asarray(vx[k], np.float32) for k in ("nx", "ny", "nz")], 1)[faces]
    return _finish(verts, faces, cuv, cn)


def load_mesh(path: str) -> Mesh:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        return load_obj(path)
    if ext == ".ply":
        return load_ply(path)
    raise ValueError(f"unsupported mesh format '{ext}' (OBJ and PLY are read)")


# ---------------------------------------------------------------------------------------------------------------
# camera (scannet_renderer.cpp:24-55, include/util.h:11-35)
# ---------------------------------------------------------------------------------------------------------------
def view_rows(pose_c2w) -> np.ndarray:
    pose = np.asarray(pose_c2w, dtype=np.float64)
    right, up, look, eye = pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3]
    right, up, look = right / np.linalg.norm(right), up / np.linalg.norm(up), look / np.linalg.norm(look)
    V = np.zeros((3, 4))
    V[0, :3], V[0, 3] = right, -right @ eye
    V[1, :3], V[1, 3] = up, -up @ eye
    V[2, :3], V[2, 3] = -look, look @ eye
    return V


def projection_entries(K, size_wh, near=NEAR, far=FAR) -> np.ndarray:
    W, H = size_wh
    K = np.asarray(K, dtype=np.float64)
    return np.array([2 * K[0, 0] / W, -(2 * (K[0, 2] / W) - 1), 2 * K[1, 1] / H, -(2 * (K[1, 2] / H) - 1),
                     -(far + near) / (far - near), -2 * far * near / (far - near)])


class MeshRasterizer:
    def __init__(self, mesh: Mesh, device=None):
        if not torch.cuda.is_available():
            raise _abi.StyleMeshB200Error("MeshRasterizer needs a CUDA device: stylemesh_b200 has no CPU path")
        self._lib = _abi.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.verts, self.faces = up(mesh.verts), up(mesh.faces)
        self.corner_uv, self.corner_normal = up(mesh.corner_uv), up(mesh.corner_normal)
        self.num_verts, self.num_faces = int(mesh.verts.shape[0]), int(mesh.faces.shape[0])
        if self.num_faces and (int(mesh.faces.min()) < 0 or int(mesh.faces.max()) >= self.num_verts):
            raise ValueError("face indices out of range")
        self._eye = torch.empty((self.num_verts, 4), device=self.device, dtype=torch.float32)
        self._zbuf: Optional[torch.Tensor] = None

    def render(self, pose_c2w, K, K_size_wh: Tuple[int, int], out_wh: Tuple[int, int], flip: bool = False,
               near: float = NEAR, far: float = FAR, tex_size: float = TEX_SIZE):
        """-> (uv, angle, depth), each (h, w, 3) float32 on the device, as main.cpp:60-67 writes them per pose."""
        import ctypes as C
        w, h = int(out_wh[0]), int(out_wh[1])
        if self._zbuf is None or self._zbuf.numel() < w * h:
            self._zbuf = torch.empty(w * h, device=self.device, dtype=torch.int64)
        outs = [torch.empty((h, w, 3), device=self.device, dtype=torch.float32) for _ in range(3)]
        V = (C.c_float * 12)(*[float(x) for x in view_rows(pose_c2w).reshape(-1)])
        P = (C.c_float * 6)(*[float(x) for x in projection_entries(K, K_size_wh, near, far)])
        _abi.check(self._lib.smb_raster_view(_abi.ptr(self.verts), self.num_verts, _abi.ptr(self.faces), self.num_faces,
                                             _abi.ptr(self.corner_uv), _abi.ptr(self.corner_normal), V, P, w, h,
                                             float(near), float(far), float(tex_size), int(bool(flip)),
                                             _abi.ptr(self._eye), _abi.ptr(self._zbuf), _abi.ptr(outs[0]),
                                             _abi.ptr(outs[1]), _abi.ptr(outs[2]), _abi.current_stream()),
                   "smb_raster_view")
        return tuple(outs)


# ---------------------------------------------------------------------------------------------------------------
# scene driver (scripts/scannet/render_uvs.py:69-107 + render_uv/src/main.cpp:60-67)
# ---------------------------------------------------------------------------------------------------------------
def read_scannet_intrinsics(path: str):
    """fx_color / fy_color / mx_color / my_color / colorWidth / colorHeight of <scene>.txt (scannet_parser.h:46-75)."""
    vals = {}
    with open(path) as fh:
        for line in fh:
            if " = " in line:
                k, v = line.split(" = ", 1)
                vals[k.strip()] = v.strip()
    K = np.eye(3)
    K[0, 0], K[1, 1] = float(vals["fx_color"]), float(vals["fy_color"])
    K[0, 2], K[1, 2] = float(vals["mx_color"]), float(vals["my_color"])
    return K, (int(vals["colorWidth"]), int(vals["colorHeight"]))


def read_pose(path: str) -> np.ndarray:
    with open(path) as fh:
        return np.array([[float(x) for x in line.split()] for line in fh if line.strip()], dtype=np.float64)


def multi_size_list(min_h=256, max_h=960, steps=5, aspect=1280 / 960) -> List[Tuple[float, int]]:
    """render_uvs.py:77-80: heights = linspace(min, max, steps) (floats: the folders are called uv_256.0 ...),
    widths = int(round(h * aspect))."""
    return [(float(h), int(round(h * aspect))) for h in np.linspace(min_h, max_h, num=steps)]


def render_scene(mesh_path: str, pose_dir: str, intrinsics_path: str, scene_dir: str, base_wh=(640, 480),
                 multi_size: Optional[Sequence[Tuple[float, int]]] = None, flip: bool = False, device=None) -> int:
    """Writes `<scene_dir>/uv/<i>.npy|.angle.npy|.rendered_depth.npy` at base_wh and, for every (height, width) of
    `multi_size`, `<scene_dir>/uv_<height>/...` - the files and names the reference's data readers expect
    (data/scannet_dataset.py:198-257; stylemesh_b200.data.scannet_scene).  Returns the number of poses rendered."""
    r = MeshRasterizer(load_mesh(mesh_path), device)
    K, K_size = read_scannet_intrinsics(intrinsics_path)
    names = sorted((f for f in os.listdir(pose_dir) if f.endswith(".txt")), key=lambda f: int(f[:-4]))
    runs = [(os.path.join(scene_dir, "uv"), (int(base_wh[0]), int(base_wh[1])))]
    for hgt, wid in (multi_size or []):
        runs.append((os.path.join(scene_dir, f"uv_{hgt}"), (int(wid), int(hgt))))
    for out_dir, wh in runs:
        os.makedirs(out_dir, exist_ok=True)
        for f in names:
            uv, ang, dep = r.render(read_pose(os.path.join(pose_dir, f)), K, K_size, wh, flip)
            stem = os.path.join(out_dir, str(int(f[:-4])))
            np.save(stem + ".npy", uv.cpu().numpy())
            np.save(stem + ".angle.npy", ang.cpu().numpy())
            np.save(stem + ".rendered_depth.npy", dep.cpu().numpy())
    return len(names)


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="CUDA replacement of scripts/scannet/render_uv + render_uvs.py for one scene")
    ap.add_argument("--mesh", required=True)
    ap.add_argument("--pose_dir", required=True)
    ap.add_argument("--intrinsics", required=True, help="ScanNet <scene>.txt")
    ap.add_argument("--out", required=True, help="scene directory (gets uv/ and uv_<h>/)")
    ap.add_argument("--w", type=int, default=640)
    ap.add_argument("--h", type=int, default=480)
    ap.add_argument("--flip", type=int, default=0)
    ap.add_argument("--multi_size", action="store_true")
    ap.add_argument("--multi_size_steps", type=int, default=5)
    ap.add_argument("--multi_size_min", type=int, default=256)
    ap.add_argument("--multi_size_max", type=int, default=960)
    ap.add_argument("--multi_size_aspect", type=float, default=1280 / 960)
    a = ap.parse_args(argv)
    ms = multi_size_list(a.multi_size_min, a.multi_size_max, a.multi_size_steps, a.multi_size_aspect) if a.multi_size else None
    n = render_scene(a.mesh, a.pose_dir, a.intrinsics, a.out, (a.w, a.h), ms, bool(a.flip))
    print(f"rendered {n} poses x {1 + len(ms or [])} sizes into {a.out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
