"""Reader for one scene in the reference's ScanNet layout (data/scannet_dataset.py:97-280) feeding a `ViewStore`,
and the DataModule behind `python -m model.optimize --dataset scannet` (model/optimize.py:44-63).

    <root_path>/train/images/<scene>/color/<i>.jpg|png        colour frames            (get_colors :97-111)
                                     depth/<i>.png            uint16 sensor depth, mm   (get_depth :113-144, load_depth :298-306)
                                     pose/<i>.txt             4x4 camera-to-world       (get_extrinsics :146-158)
                                     <scene>.txt              fx_color = ... intrinsics (get_intrinsics :160-196)
                                     uv_<height>/<i>.npy      UV pyramid level, (H,W,3) float32   (get_uvs :198-239)
                                     uv/<i>.angle.npy         cos(view angle), channel 0          (get_angles :241-257)
                                     uv/<i>.rendered_depth.npy   used when depth/ is empty        (get_depth :117-126)

File discovery and ordering follow the reference (numeric sort on the name before the first '.', duplicate `uv_256` /
`uv_256.0` folders collapsed, levels below `min_pyramid_height` dropped, the first `pyramid_levels` kept).
"""
from __future__ import annotations

import os
from os.path import isdir, join
from typing import List, Tuple

import numpy as np

from .scene_base import SceneBase, ViewStoreDataModule, load_scene_into_store  # noqa: F401  (re-exported)


def _is_float(s: str) -> bool:
    try:
        float(s)
        return True
    except ValueError:
        return False


def _numeric_sorted(folder: str, keep) -> List[str]:
    if not isdir(folder):
        return []
    files = sorted(os.listdir(folder), key=lambda x: int(x.split(".")[0]))
    return [join(folder, f) for f in files if keep(f)]


class ScanNetScene(SceneBase):
    """Paths and per-scene constants of one scene directory."""
    depth_divisor = 1000.0                 # scannet_dataset.py:301
    mask_uses_depth = True                 # scannet_dataset.py:319-320

    def __init__(self, scene_path: str, pyramid_levels: int = 5, min_pyramid_height: float = 32):
        self.path = scene_path
        if not isdir(scene_path):
            raise FileNotFoundError(f"scene directory not found: {scene_path}")
        self.colors = _numeric_sorted(join(scene_path, "color"), lambda f: f.endswith(("jpg", "png")))
        sensor = _numeric_sorted(join(scene_path, "depth"), lambda f: True)
        rendered = _numeric_sorted(join(scene_path, "uv"), lambda f: "npy" in f and "depth" in f)
        self.rendered_depth = len(sensor) == 0                     # an EMPTY depth/ folder selects the rendered depth;
        if not isdir(join(scene_path, "depth")):                   # a missing one makes the scene incomplete (:128-131)
            rendered = []
        self.depths = rendered if self.rendered_depth else sensor
        self.poses = _numeric_sorted(join(scene_path, "pose"), lambda f: True)
        self.angles = _numeric_sorted(join(scene_path, "uv"), lambda f: "npy" in f and "angle" in f)
        folders = [f for f in os.listdir(scene_path) if "uv_" in f and _is_float(f.split("_")[1])]
        folders = sorted(folders, key=lambda x: float(x.split("_")[1]))
        folders = [f for i, f in enumerate(folders)
                   if i == 0 or float(f.split("_")[1]) != float(folders[i - 1].split("_")[1])]
        self.all_levels = [float(f.split("_")[1]) for f in folders]
        folders = [f for f in folders if float(f.split("_")[1]) >= min_pyramid_height][:pyramid_levels]
        self.levels = [float(f.split("_")[1]) for f in folders]
        self.uv_levels = [_numeric_sorted(join(scene_path, f),
                                          lambda n: "npy" in n and "angle" not in n and "depth" not in n)
                          for f in folders]
        self.intrinsics, self.intrinsics_size_wh = self._read_intrinsics()
        self._check_complete()

    def _read_intrinsics(self) -> Tuple[np.ndarray, Tuple[int, int]]:
        k = np.identity(4, dtype=np.float32)
        w = h = 0
        txt = [join(self.path, f) for f in os.listdir(self.path) if ".txt" in f]
        if len(txt) == 1:
            with open(txt[0]) as fh:
                for line in fh:
                    line = line.strip()
                    for key, (r, c) in (("fx_color", (0, 0)), ("fy_color", (1, 1)), ("mx_color", (0, 2)),
                                        ("my_color", (1, 2))):
                        if key in line:
                            k[r, c] = float(line.split(" = ")[1])
                    if "colorWidth" in line:
                        w = int(line.split(" = ")[1])
                    if "colorHeight" in line:
                        h = int(line.split(" = ")[1])
        return k, (w, h)


class ScanNetViewStoreDataModule(ViewStoreDataModule):
    def __init__(self, args, device=None):
        super().__init__(args, lambda a: ScanNetScene(join(a.root_path, "train/images", a.scene),
                                                      pyramid_levels=a.pyramid_levels,
                                                      min_pyramid_height=a.min_pyramid_height), device)
