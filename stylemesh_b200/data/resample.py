"""Host-side index / weight tables of the three resampling conventions the reference's dataset uses
(data/abstract_dataset.py:299-311, data/scannet_dataset.py:319-320).  The device kernels only apply the tables
(`smb_view_gather2d`, `smb_view_resize_linear`), so the sampled indices are exact by construction:

  cv2.resize(..., INTER_LINEAR)    fx = (d + 0.5) * (src / dst) - 0.5 in float64, floor + fraction, border indices
                                   clamped with a zero weight (the IPP path of the opencv-python wheels)
  cv2.resize(..., INTER_NEAREST)   sx = min(floor(d * (src / dst)), src - 1)
  PIL Image.resize(..., NEAREST)   ImagingScaleAffine: xo = a0 / 2, xin = (int) xo, xo += a0 (accumulated in float64)
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def cv2_linear_table(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    scale = float(src) / float(dst)
    ofs = np.empty(dst, dtype=np.int32)
    alpha = np.empty(dst, dtype=np.float64)
    for d in range(dst):
        fx = (d + 0.5) * scale - 0.5
        sx = math.floor(fx)
        fx -= sx
        if sx < 0:
            sx, fx = 0, 0.0
        if sx >= src - 1:
            sx, fx = src - 1, 0.0
        ofs[d], alpha[d] = sx, fx
    return ofs, alpha


def cv2_nearest_table(src: int, dst: int) -> np.ndarray:
    scale = float(src) / float(dst)
    return np.fromiter((min(math.floor(d * scale), src - 1) for d in range(dst)), dtype=np.int32, count=dst)


def pil_nearest_table(src: int, dst: int) -> np.ndarray:
    a0 = float(src) / float(dst)
    xo = a0 * 0.5
    tab = np.empty(dst, dtype=np.int32)
    for d in range(dst):
        tab[d] = min(int(xo), src - 1)
        xo += a0
    return tab
