"""CPU oracle (numpy) of the reference's per-view input preparation — TEST INFRASTRUCTURE ONLY.

Restates, stage by stage, what `Abstract_Dataset.__getitem__` (data/abstract_dataset.py:270-344) does to the raw files
of one view with the ScanNet/Matterport implementations of its hooks (data/scannet_dataset.py:259-366,
data/matterport_dataset.py:285-311) and the transforms of model/optimize.py:33-38.  Third-party arithmetic it leans on
(not under /root/reference; versions installed in this image):
  * OpenCV 4.13 `cv2.resize` INTER_LINEAR / INTER_NEAREST on floating-point images (modules/imgproc/src/resize.cpp:
    `fx = (dx + 0.5) * scale_x - 0.5; sx = cvFloor(fx); fx -= sx`, border clamp with zero weight; double-precision
    source coordinates, arithmetic in the image's type (the IPP path of the opencv-python wheels); nearest
    `sx = min(cvFloor(dx * scale_x), w - 1)`),
  * Pillow 12.2 `Image.resize(..., NEAREST)` (src/libImaging/Geometry.c `ImagingScaleAffine`: `xo = a0 * 0.5`,
    `xin = (int) xo`, `xo += a0` accumulated in double) and `Image.resize` default BICUBIC for the colour image
    (kept on the host in the product as well: decoding and resizing the JPEG is I/O, not the hot path),
  * torchvision `ToTensor` (uint8 -> float32 `/ 255`, HWC -> CHW; float arrays are only transposed).
PINNED: tests/test_view_prep_oracle.py checks every output against tests/golden/view_prep.npz, which
tests/golden/make_view_golden.py produced by running the real reference dataset class on a synthetic scene.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(stylemesh_b200/data) never does.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

MEAN_BGR = np.array([0.40760392, 0.45795686, 0.48501961], dtype=np.float32)   # model/losses/rgb_transform.py:8


# ---------------------------------------------------------------------------------------------------------------
# resampling tables (third-party semantics, see the module docstring)
# ---------------------------------------------------------------------------------------------------------------
def cv2_linear_table(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """(ofs int32[dst], alpha float64[dst]): out = in[ofs] * (1 - alpha) + in[ofs + 1] * alpha  (cv2 INTER_LINEAR).
    The opencv-python wheels run floating-point images through IPP, whose source coordinates are double precision
    (measured here: double weights reproduce cv2 to 1 ulp on float64 and float32 images; the float weights of
    OpenCV's own C++ fallback are 2e-6 off)."""
    scale = float(src) / float(dst)
    ofs = np.zeros(dst, dtype=np.int32)
    alpha = np.zeros(dst, dtype=np.float64)
    for d in range(dst):
        fx = (d + 0.5) * scale - 0.5
        sx = int(np.floor(fx))
        fx = fx - sx
        if sx < 0:
            fx, sx = 0.0, 0
        if sx >= src - 1:
            fx, sx = 0.0, src - 1
        ofs[d], alpha[d] = sx, fx
    return ofs, alpha


def cv2_nearest_table(src: int, dst: int) -> np.ndarray:
    scale = float(src) / float(dst)            # cv2: ifx = 1 / inv_scale_x
    return np.array([min(int(np.floor(d * scale)), src - 1) for d in range(dst)], dtype=np.int32)


def pil_nearest_table(src: int, dst: int) -> np.ndarray:
    a0 = float(src) / float(dst)
    xo = a0 * 0.5
    tab = np.zeros(dst, dtype=np.int32)
    for d in range(dst):
        tab[d] = int(xo)
        xo += a0
    return np.minimum(tab, src - 1)


def resize_linear_cv2(img: np.ndarray, size_wh: Tuple[int, int]) -> np.ndarray:
    """cv2.resize(img, (w, h), interpolation=INTER_LINEAR) for a 2-D float array; work type = the input type."""
    w, h = size_wh
    src = np.asarray(img)
    if src.ndim == 3:
        src = src[:, :, 0]
    if src.shape == (h, w):
        return src.copy()
    wt = src.dtype.type if src.dtype in (np.float32, np.float64) else np.float64
    src = src.astype(wt, copy=False)
    xo, xa = cv2_linear_table(src.shape[1], w)
    yo, ya = cv2_linear_table(src.shape[0], h)
    xa, ya = xa.astype(wt), ya.astype(wt)
    x1 = np.minimum(xo + 1, src.shape[1] - 1)
    rows = src[:, xo] * (wt(1) - xa) + src[:, x1] * xa                       # horizontal pass
    y1 = np.minimum(yo + 1, src.shape[0] - 1)
    return rows[yo] * (wt(1) - ya)[:, None] + rows[y1] * ya[:, None]


def resize_nearest_cv2(img: np.ndarray, size_wh: Tuple[int, int]) -> np.ndarray:
    w, h = size_wh
    src = np.asarray(img)
    return src[cv2_nearest_table(src.shape[0], h)][:, cv2_nearest_table(src.shape[1], w)]


def resize_nearest_pil(img: np.ndarray, size_wh: Tuple[int, int]) -> np.ndarray:
    w, h = size_wh
    src = np.asarray(img)
    return src[pil_nearest_table(src.shape[0], h)][:, pil_nearest_table(src.shape[1], w)]


# ---------------------------------------------------------------------------------------------------------------
# stages
# ---------------------------------------------------------------------------------------------------------------
def uv_to_grid(uv_hw3: np.ndarray) -> np.ndarray:
    """get_uv_transform (model/texture/utils.py:87-91): ToTensor (transpose only) + to_grid = (x * 2.0) - 1 on the
    first two channels, back to HWC (utils.py:6-8, 56-60)."""
    uv = np.asarray(uv_hw3, dtype=np.float32)[:, :, :2]
    return (uv * np.float32(2.0)) - np.float32(1.0)


def uv_valid_mask(uv_hw3: np.ndarray, depth: np.ndarray = None) -> np.ndarray:
    """calculate_mask (scannet_dataset.py:308-326): (u != 0) | (v != 0), times (depth resized to the UV size > 0);
    Matterport's variant (matterport_dataset.py:295-311) ignores the depth."""
    uv = np.asarray(uv_hw3)
    m = (uv[:, :, 0] != 0) | (uv[:, :, 1] != 0)
    if depth is not None:
        d = resize_linear_cv2(depth, (m.shape[1], m.shape[0]))
        m = m & (d > 0)
    return m


def depth_levels(depth_hw: np.ndarray, levels: Sequence[float], min_depth: float):
    """calculate_depth_level (scannet_dataset.py:328-366), verbatim arithmetic in the depth's dtype (float64 for sensor
    PNGs / 1000.0)."""
    levels = np.asarray(levels, dtype=np.float64)
    n_levels = len(levels)
    depth = np.asarray(depth_hw)
    if depth.ndim == 3:
        depth = depth.squeeze()
    uv_height = 32 * (depth / min_depth)
    x = np.subtract.outer(uv_height, levels)
    rounded = np.argmin(abs(x), axis=2)
    residues = levels[rounded] - uv_height
    discrete = np.where(residues > 0, -1, 1)
    discrete[residues == 0] = 0
    other = rounded + discrete
    other[other < 0] = 0
    other[other >= n_levels] = n_levels - 1
    height_difference = abs(levels[rounded] - levels[other])
    level_residues = abs(residues / (height_difference + 1e-6))
    level_residues[height_difference == 0] = 0
    level_residues = 1 - level_residues
    continuous = np.where(residues > 0, other + level_residues, other - level_residues)
    continuous[level_residues == 1] = rounded[level_residues == 1]
    return (continuous.astype(np.float32), rounded.astype(np.int64), other.astype(np.int64),
            level_residues.astype(np.float32))


def rgb_pre(rgb_u8_hwc: np.ndarray) -> np.ndarray:
    """ToTensor + pre() (model/losses/rgb_transform.py:5-11): /255, RGB -> BGR, minus mean, times 255; float32."""
    x = np.asarray(rgb_u8_hwc, dtype=np.uint8).astype(np.float32) / np.float32(255.0)
    x = x.transpose(2, 0, 1)[[2, 1, 0]]
    x = (x - MEAN_BGR[:, None, None]) / np.float32(1.0)
    return x * np.float32(255.0)


def angle_degrees(cos_hw: np.ndarray) -> np.ndarray:
    """abstract_dataset.py:338 — torch.rad2deg(torch.acos(angle)) in float32."""
    a = np.arccos(np.asarray(cos_hw, dtype=np.float32)).astype(np.float32)
    return (a * np.float32(180.0 / np.pi)).astype(np.float32)


def modify_intrinsics(intr: np.ndarray, intr_size_wh: Tuple[int, int], rgb_size_wh: Tuple[int, int]) -> np.ndarray:
    """abstract_dataset.py:257-265."""
    k = np.array(intr, dtype=np.float32)
    if tuple(intr_size_wh) != tuple(rgb_size_wh):
        k[0, 0] = (k[0, 0] / intr_size_wh[0]) * rgb_size_wh[0]
        k[1, 1] = (k[1, 1] / intr_size_wh[1]) * rgb_size_wh[1]
        k[0, 2] = (k[0, 2] / intr_size_wh[0]) * rgb_size_wh[0]
        k[1, 2] = (k[1, 2] / intr_size_wh[1]) * rgb_size_wh[1]
    return k


def resolve_resize(resize_size, orig_wh: Tuple[int, int]) -> Tuple[int, int]:
    """abstract_dataset.py:291-297: an int is the new height, the width keeps the aspect ratio."""
    if isinstance(resize_size, int):
        w, h = orig_wh
        return int(round(w * resize_size / h)), resize_size
    return tuple(resize_size)


def preprocess_view(rgb_resized_u8: np.ndarray, uv_pyramid: List[np.ndarray], angle_hwc: np.ndarray,
                    depth_hw: np.ndarray, levels: Sequence[float], min_depth: float, size_wh: Tuple[int, int],
                    mask_uses_depth: bool = True) -> Dict[str, object]:
    """The tensors of the reference's 13-tuple that depend on pixel data (abstract_dataset.py:283-344); `rgb_resized_u8`
    is the colour image after PIL's resize, `depth_hw` the raw depth in metres (float64 for sensor depth)."""
    depth_r = resize_linear_cv2(depth_hw, size_wh)                                   # :301-304
    mask = uv_valid_mask(uv_pyramid[-1], depth_hw if mask_uses_depth else None)      # :285
    mask = resize_nearest_pil(mask, size_wh)                                         # :311
    angle = resize_nearest_cv2(np.asarray(angle_hwc)[:, :, 0], size_wh)              # :306-310
    dl, rounded, other, weight = depth_levels(depth_r, levels, min_depth)            # :317
    return {
        "rgb": rgb_pre(rgb_resized_u8),
        "depth": depth_r.astype(np.float32)[None],
        "depth_level": dl[None], "rounded_depth_level": rounded[None], "other_depth_level": other[None],
        "interp_weight": weight[None],
        "uv": [uv_to_grid(u) for u in uv_pyramid],
        "mask": mask > 0,
        "angle_guidance": angle.astype(np.float32)[None],
        "angle_degrees": angle_degrees(angle)[None],
    }
