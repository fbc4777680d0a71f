"""pre()/post() colour-space transforms with the reference's names (model/losses/rgb_transform.py:5-21).

pre():  RGB [0,1] -> BGR, minus ImageNet mean, x255  (the value range the texture lives in).
post(): the inverse plus clamp to [0,1].  Unlike the reference, neither transform mutates its input (the reference's
in-place mul_ aliases the Parameter on CPU).
"""
from __future__ import annotations

import torch

_MEAN_BGR = (0.40760392, 0.45795686, 0.48501961)


class _Transform:
    def __init__(self, fn):
        self._fn = fn

    def __call__(self, x):
        return self._fn(x)


def _mean(x):
    return torch.tensor(_MEAN_BGR, dtype=x.dtype, device=x.device).view(3, 1, 1)


def pre():
    def fwd(x):
        bgr = x[[2, 1, 0]]
        return (bgr - _mean(bgr)) * 255.0
    return _Transform(fwd)


def post():
    def inv(x):
        y = x.detach().to("cpu") * (1.0 / 255.0)
        y = y + _mean(y)
        return y[[2, 1, 0]].clamp(0, 1)
    t = _Transform(inv)
    t.is_reference_post = True      # lets the texture modules run the export chain on the device (stylemesh_b200.export)
    return t
