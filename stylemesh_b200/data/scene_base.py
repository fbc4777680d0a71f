"""Common part of the scene readers (ScanNet / Matterport layouts): decoding the files of one view into a `RawView`
(what `Abstract_Dataset.__getitem__` does before any pixel arithmetic, data/abstract_dataset.py:270-299), filling a
`ViewStore`, and the DataModule that replaces `{ScanNet,Matterport}_Single_Scene_DataModule` (model/optimize.py:44-88).
Pixel work happens on the device (ViewStore); the readers only discover / decode files and resize the colour image
with PIL, exactly as the reference does (data/abstract_dataset.py:283, 299).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from ..lightning_shim import LightningDataModule
from .view_store import RawView, ViewStore


class SceneBase:
    """Sorted file lists and per-scene constants; subclasses fill them from a directory layout."""
    path: str
    colors: List[str]
    depths: List[str]
    poses: List[str]
    angles: List[str]
    uv_levels: List[List[str]]
    levels: List[float]
    all_levels: List[float]
    rendered_depth: bool
    depth_divisor: float = 1000.0          # uint16 sensor depth -> metres
    mask_uses_depth: bool = True           # calculate_mask multiplies with depth > 0
    intrinsics: np.ndarray
    intrinsics_size_wh: Tuple[int, int]

    def __len__(self) -> int:
        return len(self.colors)

    def _check_complete(self):
        n = len(self.colors)
        ok = (n > 0 and n == len(self.depths) and len(self.uv_levels) > 0 and all(len(u) == n for u in self.uv_levels)
              and n == len(self.angles) and n == len(self.poses))
        if not ok:          # the reference silently skips such a scene (abstract_dataset.py:133-160)
            raise ValueError(f"scene {self.path} is rendered incompletely: colors {n}, depth {len(self.depths)}, "
                             f"uv {[len(u) for u in self.uv_levels]}, angles {len(self.angles)}, poses {len(self.poses)}")

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
        else:                                                                # :301 / matterport_dataset.py:288
            depth, div = np.asarray(Image.open(self.depths[i])), self.depth_divisor      # (division on the device)
            if depth.dtype != np.uint16:
                depth, div = depth.astype(np.float64) / self.depth_divisor, 1.0
        with open(self.poses[i]) as fh:                                      # load_extrinsics
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


def load_scene_into_store(scene: SceneBase, device, resize_size, min_pyramid_depth: float,
                          max_images: int = -1) -> ViewStore:
    n = len(scene) if max_images is None or max_images < 0 else min(len(scene), max_images)
    store: Optional[ViewStore] = None
    for i in range(n):
        raw, size_wh = scene.load_raw(i, resize_size)
        if store is None:
            store = ViewStore(device, scene.levels, min_pyramid_depth, size_wh, mask_uses_depth=scene.mask_uses_depth)
        store.add(raw)
    if store is None:
        raise ValueError(f"scene {scene.path} has no views")
    return store


class _RandomOrderLoader:
    """SubsetRandomSampler semantics (data/abstract_dataset.py:474-475): a fresh permutation of the indices every time
    the loader is iterated (once per epoch); torch's global generator, like the reference's sampler."""

    def __init__(self, store: ViewStore, indices):
        self.store, self.indices = store, list(indices)

    def __len__(self) -> int:
        return len(self.indices)

    def __iter__(self):
        import torch
        for j in torch.randperm(len(self.indices)).tolist():
            yield self.store[self.indices[j]]


class ViewStoreDataModule(LightningDataModule):
    """`*_Single_Scene_DataModule` (data/scannet_single_scene_dataset.py:15-64, data/matterport_single_scene_dataset.py:
    15-69) on a ViewStore: the scene is read and prepared once in setup(); the loaders hand out resident device batches
    (sampler modes 'repeat', 'sequential' and 'random'; sequential train / val split, data/abstract_dataset.py:461-470)."""

    def __init__(self, args, open_scene: Callable[[object], SceneBase], device=None):
        self.args = args
        self.device = device
        self._open_scene = open_scene
        self.train_indices: List[int] = []
        self.val_indices: List[int] = []
        self.selected_scene = ""
        self.store: Optional[ViewStore] = None

    def setup(self, stage=None):
        import torch
        a = self.args
        if not a.scene:
            raise ValueError(f"--scene is required with --dataset {a.dataset} (the reference's random scene search "
                             f"over min/max_images is not reproduced)")
        scene = self._open_scene(a)
        self.scene = scene                                  # file lists (the preview step re-reads the raw UV + LOD maps)
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
        if a.sampler_mode == "random":
            return _RandomOrderLoader(self.store, self.train_indices)      # SubsetRandomSampler(train_indices)
        raise ValueError(f"Unsupported sampler mode: {a.sampler_mode}")

    def val_dataloader(self):
        return self.store.batches(self.val_indices, 1) if self.val_indices else None
