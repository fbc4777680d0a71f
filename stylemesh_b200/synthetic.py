"""Seeded synthetic inputs with the shapes/semantics of the reference's data path (no dataset or vgg_conv.pth is
available offline).  Everything is generated on the CPU from torch.Generator seeds so that the GPU box, this
container and the golden-fixture script all see bit-identical tensors.

Mirrors: the 13-tuple of data/abstract_dataset.py:329-342 (unpacked at model/model.py:183), the Gatys VGG
state_dict keys of model/losses/content_and_style_losses.py:11-26, and the style-image preparation of
model/optimize.py:117-126 (ToTensor -> pre()).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

IMAGENET_MEAN_BGR = (0.40760392, 0.45795686, 0.48501961)

VGG_CONV_SPECS = [
    ("conv1_1", 3, 64), ("conv1_2", 64, 64),
    ("conv2_1", 64, 128), ("conv2_2", 128, 128),
    ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256),
    ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512),
    ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512),
]


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def make_vgg_state_dict(seed: int = 0, bias_scale: float = 0.0, num_convs: int = 16) -> Dict[str, torch.Tensor]:
    """He-normal N(0, 2/(9 Cin)) weights with the reference's key names (activations stay O(input) through all
    layers).  bias_scale=0 -> zero bias (benchmark config); tests use a non-zero bias."""
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, cin, cout in VGG_CONV_SPECS[:num_convs]:
        std = math.sqrt(2.0 / (9.0 * cin))
        sd[name + ".weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
        sd[name + ".bias"] = torch.randn(cout, generator=g) * bias_scale
    return sd


def pre_space(rgb01: torch.Tensor) -> torch.Tensor:
    """model/losses/rgb_transform.py:5-11 on a (3,H,W) image in [0,1] (not in place)."""
    bgr = rgb01[[2, 1, 0]].clone()
    mean = torch.tensor(IMAGENET_MEAN_BGR, dtype=bgr.dtype).view(3, 1, 1)
    return (bgr - mean) * 255.0


def make_style_image(seed: int, height: int, width: int) -> torch.Tensor:
    """A smooth coloured pattern plus noise, (3,H,W) already in pre()-space (as model/optimize.py:125-126)."""
    g = _gen(seed)
    ys = torch.linspace(0, 1, height).view(height, 1)
    xs = torch.linspace(0, 1, width).view(1, width)
    ph = torch.rand(3, 3, generator=g) * 6.283
    fr = 2.0 + 6.0 * torch.rand(3, 2, generator=g)
    chans = []
    for c in range(3):
        base = 0.5 + 0.25 * torch.sin(fr[c, 0] * 6.283 * xs + ph[c, 0]) * torch.cos(fr[c, 1] * 6.283 * ys + ph[c, 1])
        chans.append(base + 0.15 * (torch.rand(height, width, generator=g) - 0.5))
    return pre_space(torch.stack(chans).clamp(0, 1))


def make_texture_layers(seed: int, width: int, height: int, num_layers: int, random_init: bool = True,
                        channels: int = 3) -> List[torch.Tensor]:
    """model/texture/texture.py:29-32,81 — layer i is (C, H//2^i, W//2^i); rand in [0,1) or zeros."""
    g = _gen(seed)
    layers = []
    for i in range(num_layers):
        h, w = height // 2 ** i, width // 2 ** i
        layers.append(torch.rand(channels, h, w, generator=g) if random_init else torch.zeros(channels, h, w))
    return layers


@dataclass
class SyntheticView:
    """One camera view: the fields of the reference's 13-tuple that the step reads."""
    rgb: torch.Tensor                    # (1,3,Hr,Wr) pre()-space
    depth: torch.Tensor                  # (1,1,Hr,Wr)
    depth_level: torch.Tensor            # (1,1,Hr,Wr) float
    rounded_depth_level: torch.Tensor    # (1,1,Hr,Wr) int64
    other_depth_level: torch.Tensor      # (1,1,Hr,Wr) int64
    interp_weight: torch.Tensor          # (1,1,Hr,Wr) float
    uvs: List[torch.Tensor]              # each (1,H_i,W_i,2) in [-1,1]; invalid pixels exactly (-1,-1)
    mask: torch.Tensor                   # (1,Hr,Wr) bool
    angle_guidance: torch.Tensor         # (1,1,Hr,Wr) cos(theta)
    angle_degrees: torch.Tensor          # (1,1,Hr,Wr)
    index: int = 0

    def as_batch(self):
        """Order of data/abstract_dataset.py:329-342 / model/model.py:183."""
        eye = torch.eye(4).unsqueeze(0)
        return (self.rgb, eye, eye.clone(), self.depth, self.depth_level, self.rounded_depth_level,
                self.other_depth_level, self.interp_weight, torch.tensor([self.index]), self.uvs, self.mask,
                self.angle_guidance, self.angle_degrees)

    def to(self, device, non_blocking: bool = False) -> "SyntheticView":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return SyntheticView(mv(self.rgb), mv(self.depth), mv(self.depth_level), mv(self.rounded_depth_level),
                             mv(self.other_depth_level), mv(self.interp_weight), [mv(u) for u in self.uvs],
                             mv(self.mask), mv(self.angle_guidance), mv(self.angle_degrees), self.index)

    def pin(self) -> "SyntheticView":
        pm = lambda t: t.pin_memory()
        return SyntheticView(pm(self.rgb), pm(self.depth), pm(self.depth_level), pm(self.rounded_depth_level),
                             pm(self.other_depth_level), pm(self.interp_weight), [pm(u) for u in self.uvs],
                             pm(self.mask), pm(self.angle_guidance), pm(self.angle_degrees), self.index)

    def h2d_bytes(self) -> int:
        ts = [self.rgb, self.depth, self.depth_level, self.rounded_depth_level, self.other_depth_level, self.interp_weight,
              self.mask, self.angle_guidance, self.angle_degrees] + list(self.uvs)
        return int(sum(t.numel() * t.element_size() for t in ts))


def _norm_coords(h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    ys = ((torch.arange(h, dtype=torch.float32) + 0.5) / h).view(h, 1).expand(h, w)
    xs = ((torch.arange(w, dtype=torch.float32) + 0.5) / w).view(1, w).expand(h, w)
    return ys, xs


def make_view(seed: int, rgb_size: Tuple[int, int], level_sizes: Sequence[Tuple[int, int]],
              invalid_fraction: float = 0.10) -> SyntheticView:
    """Smooth UV chart into a random texture window with a contiguous invalid band (uv==(0,0) -> grid (-1,-1)),
    smooth depth levels (contiguous regions survive the 3x3 erosion), smooth cos-angle field."""
    g = _gen(seed)
    hr, wr = rgb_size
    n_levels = len(level_sizes)
    r = torch.rand(8, generator=g)
    u0, v0 = 0.05 + 0.25 * r[0].item(), 0.05 + 0.25 * r[1].item()
    su, sv = 0.35 + 0.3 * r[2].item(), 0.35 + 0.3 * r[3].item()
    ph0, ph1 = 6.283 * r[4].item(), 6.283 * r[5].item()
    band_lo = 1.0 - invalid_fraction

    uvs = []
    for (h, w) in level_sizes:
        ys, xs = _norm_coords(h, w)
        u = u0 + su * (xs + 0.03 * torch.sin(6.283 * ys + ph0))
        v = v0 + sv * (ys + 0.03 * torch.sin(6.283 * xs + ph1))
        uv = torch.stack([u.clamp(0.01, 0.99), v.clamp(0.01, 0.99)], dim=-1)
        invalid = (xs > band_lo).unsqueeze(-1)
        uv = torch.where(invalid, torch.zeros_like(uv), uv)
        uvs.append((uv * 2.0 - 1).unsqueeze(0).contiguous())          # to_grid_range, utils.py:6-8

    ys, xs = _norm_coords(hr, wr)
    mask = (xs <= band_lo).unsqueeze(0)
    rgb = pre_space(torch.rand(3, hr, wr, generator=g)).unsqueeze(0)
    cosang = (0.6 + 0.4 * torch.sin(3.0 * xs + 2.0 * ys + ph0)).clamp(0.2, 1.0).view(1, 1, hr, wr)
    angle_deg = torch.rad2deg(torch.acos(cosang.clamp(-1, 1)))
    depth = (0.5 + 3.5 * (0.5 + 0.5 * torch.sin(2.2 * xs + 1.3 * ys + ph1))).view(1, 1, hr, wr)
    if n_levels > 1:
        lvl = (n_levels - 1) * (0.5 + 0.5 * torch.sin(2.5 * xs - 1.7 * ys + ph0 + ph1))
    else:
        lvl = torch.zeros(hr, wr)
    lvl = lvl.view(1, 1, hr, wr)
    rounded = torch.round(lvl).clamp(0, n_levels - 1)
    other = torch.where(lvl >= rounded, rounded + 1, rounded - 1).clamp(0, n_levels - 1)
    w_interp = (1.0 - (lvl - rounded).abs()).clamp(0, 1)
    return SyntheticView(rgb=rgb, depth=depth, depth_level=lvl, rounded_depth_level=rounded.long(),
                         other_depth_level=other.long(), interp_weight=w_interp, uvs=uvs, mask=mask,
                         angle_guidance=cosang, angle_degrees=angle_deg, index=seed)


# flag families of scripts/train/optimize_texture_*.sh (loss weights and modes)
PRESETS = {
    "only2D": dict(use_angle_weight=False, use_depth_scaling=False, style_pyramid_mode="single",
                   gram_mode="current", angle_threshold=3000.0, pyramid_levels=1,
                   loss_weights={"content": 70.0, "style": 1e-4, "tex_reg": 5e3},
                   style_weights=[1000.0, 1000.0, 10.0, 10.0, 1000.0], hierarchical_layers=4),
    "with_angle": dict(use_angle_weight=True, use_depth_scaling=False, style_pyramid_mode="multi",
                       gram_mode="current", angle_threshold=30.0, pyramid_levels=1,
                       loss_weights={"content": 70.0, "style": 1e-4, "tex_reg": 5e3},
                       style_weights=[1000.0, 1000.0, 10.0, 10.0, 1000.0], hierarchical_layers=4),
    "with_angle_and_depth": dict(use_angle_weight=True, use_depth_scaling=True, style_pyramid_mode="multi",
                                 gram_mode="current", angle_threshold=30.0, pyramid_levels=4,
                                 loss_weights={"content": 70.0, "style": 1e-4, "tex_reg": 5e3},
                                 style_weights=[1000.0, 1000.0, 10.0, 10.0, 1000.0], hierarchical_layers=4),
    "dip": dict(use_angle_weight=False, use_depth_scaling=False, style_pyramid_mode="single",
                gram_mode="average", angle_threshold=3000.0, pyramid_levels=1,
                loss_weights={"content": 70.0, "style": 1e-3, "tex_reg": 0.0},
                style_weights=[1000.0, 1000.0, 10.0, 10.0, 1000.0], hierarchical_layers=1),
    "content_only": dict(use_angle_weight=False, use_depth_scaling=False, style_pyramid_mode="single",
                         gram_mode="current", angle_threshold=3000.0, pyramid_levels=1,
                         loss_weights={"content": 70.0, "style": 0.0, "tex_reg": 0.0},
                         style_weights=[1000.0, 1000.0, 10.0, 10.0, 1000.0], hierarchical_layers=1),
}


# the reference's rendered UV pyramids (scripts/scannet/render_uvs.py:77-90,126-130: heights linspace(256, 784, 4),
# widths from the sensor aspect ratio, rounded by the renderer driver) - SURVEY §8
SCANNET_PYRAMID = [(256, 341), (432, 576), (608, 811), (784, 1045)]
MATTERPORT_PYRAMID = [(256, 320), (432, 540), (608, 760), (784, 980)]


def pyramid_sizes(base_hw: Tuple[int, int], levels: int, min_height: int = None) -> List[Tuple[int, int]]:
    """UV pyramid sizes of scripts/scannet/render_uvs.py:77-90,126-130: the reference's own ScanNet / Matterport
    tables when the base size is theirs, else heights linearly spaced from the base height to ~3x with widths scaled
    by the aspect ratio."""
    h0, w0 = base_hw
    if levels <= 1:
        return [(h0, w0)]
    for table in (SCANNET_PYRAMID, MATTERPORT_PYRAMID):
        if (h0, w0) == table[0] and levels <= len(table):
            return list(table[:levels])
    out = []
    for i in range(levels):
        h = int(round(h0 + (h0 * 2.0625) * i / (levels - 1)))
        out.append((h, int(h * w0 / h0)))
    return out
