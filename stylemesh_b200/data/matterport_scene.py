"""Reader for one region of a Matterport3D house in the reference's layout (data/matterport_dataset.py:96-255) and
the DataModule behind `--dataset matterport` (model/optimize.py:65-88).

    <root_path>/v1/scans/<house>/rendered/region_<k>/color/<hash>_i<c>_<y>.jpg           (get_colors :100-113)
                                                     depth/<hash>_d<c>_<y>.png           uint16, 1/4000 m (load_depth :285-293)
                                                     rendered_depth/*.rendered_depth.npy  used when depth/ is missing or empty
                                                     pose/<hash>_i<c>_<y>.txt             4x4 camera-to-world (get_extrinsics :148-160)
                                                     pose/<name>.intrinsics.txt           3 rows of K + "W H" (get_intrinsics :162-189)
                                                     uv_<w>_<h>/<hash>_i<c>_<y>.*uvs*.npy  UV pyramid level (get_uvs :191-224)
                                                     angle/<hash>_i<c>_<y>.*angle*.npy    cos(view angle) (get_angles :226-243)

Files are ordered by (hash, camera * 100 + yaw) of the name before the first '.' (sort_keys["default"] :59-63); pyramid
folders by their last `_`-separated integer (the height).  The validity mask ignores the depth
(calculate_mask :295-311) and the sensor depth is in 1/4000 m.
"""
from __future__ import annotations

import os
from os.path import isdir, join
from typing import List, Tuple

import numpy as np

from .scene_base import SceneBase, ViewStoreDataModule


def _key(name: str):
    stem = name.split(".")[0].split("_")
    return [stem[0], int(stem[1][1]) * 100 + int(stem[2])]


def _sorted(folder: str, keep) -> List[str]:
    if not isdir(folder):
        return []
    return [join(folder, f) for f in sorted(os.listdir(folder), key=_key) if keep(f)]


class MatterportRegion(SceneBase):
    depth_divisor = 4000.0                 # matterport_dataset.py:288
    mask_uses_depth = False                # matterport_dataset.py:304-307

    def __init__(self, house_path: str, region_index: int = 0, pyramid_levels: int = 5, min_pyramid_height: int = 256):
        self.path = join(house_path, "rendered", f"region_{region_index}")
        if not isdir(self.path):
            raise FileNotFoundError(f"region directory not found: {self.path}")
        self.colors = _sorted(join(self.path, "color"), lambda f: f.endswith(("jpg", "png")))
        sensor = _sorted(join(self.path, "depth"), lambda f: True)
        rendered = _sorted(join(self.path, "rendered_depth"), lambda f: "npy" in f and "depth" in f)
        self.rendered_depth = len(sensor) == 0
        self.depths = rendered if self.rendered_depth else sensor
        self.poses = _sorted(join(self.path, "pose"), lambda f: "intrinsic" not in f)
        self.angles = _sorted(join(self.path, "angle"), lambda f: "npy" in f and "angle" in f)
        folders = sorted([f for f in os.listdir(self.path) if "uv_" in f], key=lambda x: int(x.split("_")[-1]))
        self.all_levels = [float(int(f.split("_")[-1])) for f in folders]
        folders = [f for f in folders if int(f.split("_")[-1]) >= min_pyramid_height][:pyramid_levels]
        self.levels = [float(f.split("_")[-1]) for f in folders]
        self.uv_levels = [_sorted(join(self.path, f), lambda n: "npy" in n and "uvs" in n) for f in folders]
        self.intrinsics, self.intrinsics_size_wh = self._read_intrinsics()
        self._check_complete()

    def _read_intrinsics(self) -> Tuple[np.ndarray, Tuple[int, int]]:
        k = np.identity(4, dtype=np.float32)
        w = h = 0
        pose_dir = join(self.path, "pose")
        files = [join(pose_dir, f) for f in os.listdir(pose_dir) if ".intrinsics.txt" in f] if isdir(pose_dir) else []
        if files:
            with open(files[0]) as fh:
                for i, line in enumerate(fh.readlines()):
                    e = line.strip().split(" ")
                    if i < 3:
                        k[i][0], k[i][1], k[i][2] = float(e[0]), float(e[1]), float(e[2])
                    elif i == 3:
                        w, h = int(e[0]), int(e[1])
                    else:
                        raise ValueError("index too large", i, line)
        return k, (w, h)


class MatterportViewStoreDataModule(ViewStoreDataModule):
    def __init__(self, args, device=None):
        super().__init__(args, lambda a: MatterportRegion(join(a.root_path, "v1/scans", a.scene),
                                                          region_index=a.matterport_region_index,
                                                          pyramid_levels=a.pyramid_levels,
                                                          min_pyramid_height=a.min_pyramid_height), device)
