"""UV / image tensor helpers with the names of the reference's model/texture/utils.py (drop-in surface).

Semantics restated from model/texture/utils.py:6-91: UV images are CHW in [0,1]; grid_sample grids are HWC in
[-1,1] with only the first two channels.
"""
from __future__ import annotations

import torch


def to_grid_range(x):
    """[0,1] -> [-1,1] (utils.py:6-8); fp32 op order (x*2)-1 is part of the bit-exact UV contract."""
    return (x * 2.0) - 1


def from_grid_range(x):
    """[-1,1] -> [0,1] (utils.py:11-13)."""
    return (x + 1) / 2.0


def cut_b_channel(x):
    """keep the (u,v) channels of a CHW UV image (utils.py:16-18)."""
    return x[:2]


def add_b_channel(x):
    """append a constant -1 third channel to a 2-channel CHW tensor (utils.py:21-23)."""
    filler = torch.full_like(x[:1], -1)
    return torch.cat((x, filler), dim=0)


def chw_to_hwc(x):
    """CHW -> HWC, or BCHW -> BHWC (utils.py:26-31)."""
    return x.permute(1, 2, 0) if x.dim() == 3 else x.permute(0, 2, 3, 1)


def hwc_to_chw(x):
    """HWC -> CHW, or BHWC -> BCHW (utils.py:34-39)."""
    return x.permute(2, 0, 1) if x.dim() == 3 else x.permute(0, 3, 1, 2)


def to_grid_format(x):
    return chw_to_hwc(cut_b_channel(x))


def from_grid_format(x):
    return add_b_channel(hwc_to_chw(x))


def to_grid(x):
    """CHW UV image in [0,1] -> grid_sample grid (utils.py:56-60)."""
    return to_grid_format(to_grid_range(x))


def from_grid(x):
    """inverse of to_grid (utils.py:63-67)."""
    return from_grid_range(from_grid_format(x))


def numpy_to_pil(x):
    from PIL import Image
    return Image.fromarray(x)


class _Pipeline:
    """minimal stand-in for torchvision.transforms.Compose (callable chain)."""

    def __init__(self, *fns):
        self.fns = fns

    def __call__(self, x):
        for fn in self.fns:
            x = fn(x)
        return x


def _to_tensor(pic):
    from torchvision.transforms.functional import to_tensor
    return pic if isinstance(pic, torch.Tensor) else to_tensor(pic)


def get_rgb_transform():
    return _Pipeline(_to_tensor)


def get_label_transform():
    return _Pipeline(_to_tensor)


def get_uv_transform():
    return _Pipeline(_to_tensor, to_grid)
