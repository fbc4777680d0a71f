"""Reader for one scene in the reference's ScanNet layout (data/scannet_dataset.py:97-280) feeding a `ViewStore`,
and the DataModule behind `python -m model.optimize --dataset scannet` (model/optimize.py:44-63).

    <root_path>/train/images/<scene>/color/<i>.jpg|png        colour frames            (get_colors :97-111)
                                     depth/<i>.png            uint16 sensor depth, mm   (get_depth :113-144, load_depth :298-306)
                                     pose/<i>.txt             4x4 camera-to-world       (get_extrinsics :146-158, load_extrinsics :259-272)
                                     <scene>.txt              fx_color = ... intrinsics (get_intrinsics :160-196)
                                     uv_<height>/<i>.npy      UV pyramid level, (H,W,3) float32   (get_uvs :198-239)
                                     uv/<i>.angle.npy         cos(view angle), channel 0          (get_angles :241-257)
                                     uv/<i>.rendered_depth.npy   used when depth/ is empty        (get_depth :117-126)

File discovery and ordering follow the reference (numeric sort on the name before the first '.', duplicate `uv_256` /
`uv_256.0` folders collapsed, levels below `min_pyramid_height` dropped, the first `pyramid_levels` kept).  Pixel work
happens on the device (ViewStore); this module only decodes files and resizes the colour image with PIL, exactly as
`Abstract_Dataset.__getitem__` does (data/abstract_dataset.py:283, 299).
"""
from __future__ import annotations

import os
from os.path import isdir, join
from typing import List, Optional, Tuple

import numpy as np

from ..lightning_shim import LightningDataModule
from .view_store import RawView, ViewStore


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


class ScanNetScene:
    """Paths and per-scene constants of one scene directory."""

    def __init__(self, scene_path: str, pyramid_levels: int = 5, min_pyramid_height: float = 32):
        self.path = scene_path
        if not isdir(scene_path):
            raise FileNotFoundError(f"scene directory not found: {scene_path}")
        self.colors = _numeric_sorted(join(scene_path, "color"), lambda f: f.endswith(("jpg", "png")))
        sensor = _numeric_sorted(join(scene_path, "depth"), lambda f: True)
        rendered = _numeric_sorted(join(scene_path, "uv"), lambda f: "npy" in f and "depth" in f)
        self.rendered_depth = len(sensor) == 0
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
        n = len(self.colors)
        ok = (n > 0 and n == len(self.depths) and len(self.uv_levels) > 0 and all(len(u) == n for u in self.uv_levels)
              and n == len(self.angles) and n == len(self.poses))
        if not ok:          # the reference silently skips such a scene (abstract_dataset.py:133-160)
            raise ValueError(f"scene {scene_path} is rendered incompletely: colors {n}, depth {len(self.depths)}, "
                             f"uv {[len(u) for u in self.uv_levels]}, angles {len(self.angles)}, poses {len(self.poses)}")

    def __len__(self) -> int:
        return len(self.colors)

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

    def load_raw(self, i: int, resize_size) -> Tuple[RawView, Tuple[int, int]]:
        """Decode the files of view i; returns the RawView and the working (width, height)."""
        from PIL import Image
        rgb = Image.open(self.colors[i])
        if isinstance(resize_size, int):                                     # abstract_dataset.py:291-297
            w, h = rgb.size
            size_wh = (round(w * resize_size / h), resize_size)
        else:
            size_wh = tuple(resize_size)
        rgb = np.asarray(rgb.convert("RGB").resize(size_wh))                 # :299, PIL's default filter
        if self.rendered_depth:
            depth, div = np.load(self.depths[i])[:, :, :1], 1.0              # scannet_dataset.py:303-304
        else:
            depth, div = np.asarray(Image.open(self.depths[i])), 1000.0      # :301 (the division runs on the device)
            if depth.dtype != np.uint16:
                depth, div = depth.astype(np.float64) / 1000.0, 1.0
        with open(self.poses[i]) as fh:                                      # :259-272
            extr = np.array([[float(v) for v in line.split(" ")] for line in fh.readlines()], dtype=np.float32)
        intr = np.array(self.intrinsics)                                     # abstract_dataset.py:257-265
        iw, ih = self.intrinsics_size_wh
        if (iw, ih) != size_wh:
            intr[0, 0] = (intr[0, 0] / iw) * size_wh[0]
            intr[1, 1] = (intr[1, 1] / ih) * size_wh[1]
            intr[0, 2] = (intr[0, 2] / iw) * size_wh[0]
            intr[1, 2] = (intr[1, 2] / ih) * size_wh[1]
        raw = RawView(rgb=rgb, uv_pyramid=[np.load(level[i]) for level in self.uv_levels],
                      angle=np.load(self.angles[i])[:, :, :1], depth=depth, depth_divisor=div, extrinsics=extr,
                      intrinsics=intr, index=i)
        return raw, size_wh


def load_scene_into_store(scene: ScanNetScene, device, resize_size, min_pyramid_depth: float,
                          max_images: int = -1) -> ViewStore:
    n = len(scene) if max_images is None or max_images < 0 else min(len(scene), max_images)
    store: Optional[ViewStore] = None
    for i in range(n):
        raw, size_wh = scene.load_raw(i, resize_size)
        if store is None:
            store = ViewStore(device, scene.levels, min_pyramid_depth, size_wh, mask_uses_depth=True)
        store.add(raw)
    if store is None:
        raise ValueError(f"scene {scene.path} has no views")
    return store


class ScanNetViewStoreDataModule(LightningDataModule):
    """`ScanNet_Single_Scene_DataModule` (data/scannet_single_scene_dataset.py:15-64) on a ViewStore: the scene is read
    and prepared once in setup(); the loaders hand out resident device batches (sampler modes 'repeat' and
    'sequential'; sequential train / val split, data/abstract_dataset.py:461-470)."""

    def __init__(self, args, device=None):
        self.args = args
        self.device = device
        self.train_indices: List[int] = []
        self.val_indices: List[int] = []
        self.selected_scene = ""
        self.store: Optional[ViewStore] = None

    def setup(self, stage=None):
        import torch
        a = self.args
        if not a.scene:
            raise ValueError("--scene is required with --dataset scannet (the reference's random scene search over "
                             "min/max_images is not reproduced)")
        scene = ScanNetScene(join(a.root_path, "train/images", a.scene), pyramid_levels=a.pyramid_levels,
                             min_pyramid_height=a.min_pyramid_height)
        n = len(scene)
        if not ((a.min_images == -1 or n >= a.min_images) and (a.max_images == -1 or n <= a.max_images)):
            raise ValueError(f"scene {a.scene} has {n} images, outside [--min_images {a.min_images}, "
                             f"--max_images {a.max_images}]")          # get_scene / in_range, single_scene:104-120
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        self.store = load_scene_into_store(scene, dev, a.resize_size, a.min_pyramid_depth)
        self.selected_scene = a.scene
        indices = list(range(n))
        if getattr(a, "shuffle", False):
            np.random.shuffle(indices)
        n_train = int(a.train_split * n)
        self.train_indices, self.val_indices = indices[:n_train], indices[n_train:]

    def train_dataloader(self):
        a = self.args
        if a.sampler_mode == "repeat":
            return self.store.batches(self.train_indices, a.index_repeat)
        if a.sampler_mode == "sequential":
            return self.store.batches(range(len(self.store)), 1)           # SequentialSampler(train_dataset)
        raise ValueError(f"Unsupported sampler mode: {a.sampler_mode} ('random' needs a per-epoch permutation: use "
                         f"--shuffle with 'repeat' or 'sequential')")

    def val_dataloader(self):
        return self.store.batches(self.val_indices, 1) if self.val_indices else None
